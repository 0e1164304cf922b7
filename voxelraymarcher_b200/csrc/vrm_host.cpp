// vrm_host.cpp -- host-side helpers of the C ABI that must reproduce the reference's HOST arithmetic bit for bit
// (compiled with -ffp-contract=off; same libm tanf as a host build of the reference).
#include "../../include/vrm_b200.h"

#include <math.h>

namespace
{
const float kPi = 3.141592f;  // math/MathConstants.cuh:3 (sic)

void unit(const float* v, float* out)  // makeUnitVector, math/Vector3.cuh:161-165: v / length(), IEEE division
{
	float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	out[0] = v[0] / l; out[1] = v[1] / l; out[2] = v[2] / l;
}
void cross(const float* a, const float* b, float* out)  // math/Vector3.cuh:153-158
{
	out[0] = a[1] * b[2] - a[2] * b[1];
	out[1] = -(a[0] * b[2] - a[2] * b[0]);
	out[2] = a[0] * b[1] - a[1] * b[0];
}
}  // namespace

extern "C" int vrm_make_unit_vector(const float v[3], float out[3])
{
	if (!v || !out) return VRM_ERR_INVALID;
	unit(v, out);
	return VRM_OK;
}

// Camera::Camera, renderer/camera/Camera.cuh:11-23
extern "C" int vrm_camera_make(const float origin[3], const float look_at[3], const float up[3], float fov_degrees, float aspect,
                               float out_camera[VRM_CAMERA_FLOATS])
{
	if (!origin || !look_at || !up || !out_camera) return VRM_ERR_INVALID;
	float halfHeight = tanf((fov_degrees * kPi / 180.f) / 2.0f);
	float halfWidth = halfHeight * aspect;
	float toTarget[3] = {look_at[0] - origin[0], look_at[1] - origin[1], look_at[2] - origin[2]};
	float w[3], wxup[3], u[3], v[3];
	unit(toTarget, w);
	cross(w, up, wxup);
	unit(wxup, u);
	cross(u, w, v);
	for (int i = 0; i < 3; i++)
	{
		out_camera[i] = origin[i];
		out_camera[3 + i] = origin[i] - halfWidth * u[i] - halfHeight * v[i] + w[i];  // lowerLeftCorner
		out_camera[6 + i] = (2 * halfWidth) * u[i];                                    // horizontalVector
		out_camera[9 + i] = (2 * halfHeight) * v[i];                                   // verticalVector
		out_camera[12 + i] = w[i];                                                     // forwardVector
	}
	return VRM_OK;
}
