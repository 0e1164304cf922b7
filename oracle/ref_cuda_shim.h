// TEST INFRASTRUCTURE ONLY.  Force-included (-include) before the UNMODIFIED reference headers when they are
// compiled with nvcc on Linux: the reference was written for MSVC and relies on std::floorf / std::ceilf and on
// <cstring> being pulled in transitively (SURVEY.md §8c "GPU oracle / baseline").
#pragma once
#include <cstring>
#include <cmath>
#include <math.h>
#include <string>
#include <vector>
#include <cuda_runtime.h>
namespace std { using ::floorf; using ::ceilf; }
