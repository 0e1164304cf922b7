// VoxelRaymarcher CLI over the C ABI of libvrm_b200.so -- the reference's process-level seam:
//   VoxelRaymarcher <scale:int> <hashtable|vcs> <original|longestaxis>        (main/Main.cu:176-229)
// Same positional arguments, defaults (anything but the exact strings selects VCS / Longest Axis), prints, fixed
// 1920x1080 frame and camera, `resources/scene.vox` input and `output.png` output as the reference.  Extras (all optional,
// after the positional arguments): --scene PATH --out PATH --width W --height H.
//
// Scene file: the reference's text format, one `x,y,z,color` line per voxel (geometry/VoxelFile.cuh:9-35); files that
// start with the magic "VOX " are read as MagicaVoxel binaries instead (north star; the reference itself has no such
// reader, so only the text path has reference semantics to match).
#include "../../include/vrm_b200.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace
{

// ---- scene readers ---------------------------------------------------------------------------------------------
bool readCsvScene(std::istream& in, std::vector<int32_t>& xyz, std::vector<uint32_t>& rgb)
{
	std::string line;
	while (in)
	{
		std::vector<std::string> entries;
		size_t start = 0, end = 0;
		std::getline(in, line);
		// split on commas, empty fields skipped (VoxelFile.cuh:20-24)
		while ((start = line.find_first_not_of(",", end)) != std::string::npos)
		{
			end = line.find_first_of(",", start);
			entries.push_back(line.substr(start, end - start));
		}
		if (entries.size() > 3)  // VoxelFile.cuh:27-34
		{
			xyz.push_back(std::stoi(entries[0]));
			xyz.push_back(std::stoi(entries[1]));
			xyz.push_back(std::stoi(entries[2]));
			rgb.push_back(static_cast<uint32_t>(std::stoi(entries[3])));
		}
	}
	return true;
}

uint32_t rd32(const std::vector<uint8_t>& b, size_t o) { return b[o] | (b[o + 1] << 8) | (b[o + 2] << 16) | (uint32_t(b[o + 3]) << 24); }

// MagicaVoxel .vox (version 150/200): MAIN > (SIZE, XYZI)*, optional RGBA palette.  MagicaVoxel is z-up; the
// raymarcher is y-up, so (x, y, z)_vox -> (x, z, y).  Models are placed at the origin (scene-graph nodes ignored).
bool readMagicaVoxel(const std::vector<uint8_t>& b, std::vector<int32_t>& xyz, std::vector<uint32_t>& rgb)
{
	if (b.size() < 20 || memcmp(b.data(), "VOX ", 4) != 0 || memcmp(b.data() + 8, "MAIN", 4) != 0) return false;
	std::vector<uint32_t> palette(256);
	for (int i = 0; i < 256; i++) palette[i] = 0x010101u * uint32_t(i);  // grey ramp unless an RGBA chunk follows
	std::vector<uint8_t> colourIndex;
	size_t pos = 20;
	while (pos + 12 <= b.size())
	{
		const char* id = reinterpret_cast<const char*>(b.data() + pos);
		uint32_t n = rd32(b, pos + 4), m = rd32(b, pos + 8);
		size_t body = pos + 12;
		if (body + n > b.size()) break;
		if (memcmp(id, "XYZI", 4) == 0 && n >= 4)
		{
			uint32_t count = rd32(b, body);
			for (uint32_t i = 0; i < count && body + 8 + 4 * size_t(i) <= b.size(); i++)
			{
				const uint8_t* v = b.data() + body + 4 + 4 * size_t(i);
				xyz.push_back(v[0]); xyz.push_back(v[2]); xyz.push_back(v[1]);
				colourIndex.push_back(v[3]);
			}
		}
		else if (memcmp(id, "RGBA", 4) == 0 && n >= 1024)
		{
			for (int i = 0; i < 255; i++)  // palette entry i is stored at index i+1
			{
				const uint8_t* c = b.data() + body + 4 * size_t(i);
				palette[i + 1] = (uint32_t(c[0]) << 16) | (uint32_t(c[1]) << 8) | c[2];
			}
		}
		pos = body + n;  // children of MAIN follow in line (m is the size of its children, not skipped)
		(void)m;
	}
	for (uint8_t ci : colourIndex) rgb.push_back(palette[ci]);
	return true;
}

// ---- PNG writer (8-bit RGB, stored deflate blocks; stands in for stbi_write_png, images/ImageWriter.cpp:8-16) ------------
uint32_t crcTable[256];
void initCrc()
{
	for (uint32_t n = 0; n < 256; n++)
	{
		uint32_t c = n;
		for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
		crcTable[n] = c;
	}
}
uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu)
{
	for (size_t i = 0; i < n; i++) c = crcTable[(c ^ p[i]) & 0xFF] ^ (c >> 8);
	return c;
}
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data)
{
	put32(out, uint32_t(data.size()));
	size_t at = out.size();
	out.insert(out.end(), type, type + 4);
	out.insert(out.end(), data.begin(), data.end());
	put32(out, crc32(out.data() + at, out.size() - at) ^ 0xFFFFFFFFu);
}
bool writePng(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h)
{
	initCrc();
	std::vector<uint8_t> raw;
	raw.reserve(size_t(h) * (size_t(w) * 3 + 1));
	for (uint32_t y = 0; y < h; y++)
	{
		raw.push_back(0);  // filter: none
		raw.insert(raw.end(), rgb + size_t(y) * w * 3, rgb + size_t(y + 1) * w * 3);
	}
	std::vector<uint8_t> z = {0x78, 0x01};
	uint32_t a = 1, b = 0;
	for (size_t off = 0; off < raw.size() || off == 0; off += 65535)
	{
		size_t n = std::min<size_t>(65535, raw.size() - off);
		bool last = off + n >= raw.size();
		z.push_back(last ? 1 : 0);
		z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
		z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
		for (size_t i = 0; i < n; i++) { a = (a + raw[off + i]) % 65521; b = (b + a) % 65521; }
		if (last) break;
	}
	put32(z, (b << 16) | a);
	std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	std::vector<uint8_t> ihdr;
	put32(ihdr, w); put32(ihdr, h);
	ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
	chunk(out, "IHDR", ihdr);
	chunk(out, "IDAT", z);
	chunk(out, "IEND", {});
	std::ofstream f(path, std::ios::binary);
	f.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out.size()));
	return bool(f);
}

}  // namespace

int main(int argc, char* argv[])
{
	if (argc <= 1)  // Main.cu:181-185
	{
		std::cout << "You need to provide a voxel scale" << std::endl;
		return 1;
	}
	int32_t scale = std::stoi(argv[1]);
	int storage = VRM_STORAGE_VCS, algorithm = VRM_ALGO_LONGEST_AXIS;
	if (argc > 2 && std::strcmp(argv[2], "hashtable") == 0) { std::cout << "Storage Type: Cuckoo Hash Table" << std::endl; storage = VRM_STORAGE_HASHTABLE; }  // Main.cu:45-55
	else std::cout << "Storage Type: Voxel Cluster Storage" << std::endl;
	if (argc > 3 && std::strcmp(argv[3], "original") == 0) { std::cout << "Raymarching Algorithm: Original" << std::endl; algorithm = VRM_ALGO_ORIGINAL; }  // Main.cu:58-68
	else std::cout << "Raymarching Algorithm: Longest Axis" << std::endl;
	std::string scenePath = "resources/scene.vox", outPath = "output.png";  // Main.cu:99,170; VoxelFile.cuh:12
	uint32_t width = 1920, height = 1080;                                   // Main.cu:195-196
	for (int i = 4; i + 1 < argc; i += 2)
	{
		if (!std::strcmp(argv[i], "--scene")) scenePath = argv[i + 1];
		else if (!std::strcmp(argv[i], "--out")) outPath = argv[i + 1];
		else if (!std::strcmp(argv[i], "--width")) width = uint32_t(std::stoul(argv[i + 1]));
		else if (!std::strcmp(argv[i], "--height")) height = uint32_t(std::stoul(argv[i + 1]));
	}

	char name[256] = "";
	int devCount = vrm_device_count();  // pickCudaDevice, Main.cu:82-94
	printf("Device Count: %d\n", devCount);
	if (devCount) vrm_device_name(0, name, sizeof(name));
	printf("Device: %s\n", name);

	vrm_scene* scene = nullptr;
	int rc = vrm_scene_create(0, &scene);
	if (rc != VRM_OK) { std::cout << "ERROR: " << vrm_error_string(rc) << " (a CUDA device is required)" << std::endl; return EXIT_FAILURE; }

	std::vector<int32_t> xyz;
	std::vector<uint32_t> rgb;
	{
		std::ifstream f(scenePath, std::ios::binary);
		std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
		if (bytes.size() >= 4 && memcmp(bytes.data(), "VOX ", 4) == 0) readMagicaVoxel(bytes, xyz, rgb);
		else
		{
			std::ifstream text(scenePath);
			readCsvScene(text, xyz, rgb);  // a missing file yields an empty scene, as in the reference
		}
	}
	rc = vrm_scene_add_voxels(scene, xyz.data(), rgb.data(), rgb.size());
	float buildMs = 0.0f;
	if (rc == VRM_OK) rc = vrm_scene_build(scene, storage, &buildMs);
	if (rc != VRM_OK) { std::cout << "ERROR: " << vrm_error_string(rc) << ": " << vrm_last_error(scene) << std::endl; return EXIT_FAILURE; }
	uint32_t diameter = 0, filled = 0;
	int32_t minCoord = 0;
	vrm_scene_info(scene, &diameter, &minCoord, &filled, nullptr, nullptr);
	std::cout << "There are : " << filled << "/" << diameter * diameter * diameter << " regions that are filled" << std::endl;  // VoxelSceneCPU.cuh:54
	std::cout << "Storage Structures Generated" << std::endl;                                                                  // VoxelSceneCPU.cuh:86

	float origin[3] = {6.0f, 2.0f, 6.0f}, lookAt[3] = {0.0f, 0.0f, -1.0f}, up[3] = {0.0f, 1.0f, 0.0f}, cam[VRM_CAMERA_FLOATS];  // Main.cu:199
	float aspect = static_cast<float>(width) / static_cast<float>(height);
	vrm_camera_make(origin, lookAt, up, 60.0f, aspect, cam);
	float translation[3] = {0.0f, 0.0f, 0.0f}, kernelMs = 0.0f;  // Main.cu:215
	std::vector<uint8_t> frame(size_t(width) * height * 3);
	rc = vrm_render(scene, cam, translation, uint32_t(scale), algorithm, width, height, frame.data(), nullptr, &kernelMs);
	std::cout << (rc == VRM_OK ? "no error" : vrm_last_error(scene)) << std::endl;  // cudaGetErrorString(cudaPeekAtLastError()), Main.cu:154-155
	std::cout << "Execution Time for Ray Marching Algorithm is: " << static_cast<long long>(kernelMs * 1000.0f) << " microseconds" << std::endl;  // Main.cu:162
	if (rc == VRM_OK && !writePng(outPath, frame.data(), width, height)) std::cout << "ERROR: Failed to write image to: " << outPath << std::endl;  // ImageWriter.cpp:11-14
	vrm_scene_destroy(scene);
	return rc == VRM_OK ? EXIT_SUCCESS : EXIT_FAILURE;
}
