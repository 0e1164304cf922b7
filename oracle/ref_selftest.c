/* TEST INFRASTRUCTURE ONLY.  Loads a freshly linked oracle/_ref/libvrm_ref_host.so and runs one small build + lookup of both
 * storage structures, so that oracle/Makefile can reject a host compiler whose build of the reference does not work before
 * the tests trust the library (the Makefile separately rejects a statically linked libstdc++). */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>

typedef void* (*create_fn)(void);
typedef void (*add_fn)(void*, const int32_t*, const uint32_t*, uint64_t);
typedef int (*build_fn)(void*, int);
typedef int (*lookup_fn)(void*, const int32_t*, uint64_t, uint32_t*, uint8_t*);

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	void* lib = dlopen(argv[1], RTLD_NOW);
	if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 3; }
	create_fn create = (create_fn)dlsym(lib, "refh_scene_create");
	add_fn add = (add_fn)dlsym(lib, "refh_scene_add_voxels");
	build_fn build = (build_fn)dlsym(lib, "refh_scene_build");
	lookup_fn lookup = (lookup_fn)dlsym(lib, "refh_lookup");
	if (!create || !add || !build || !lookup) return 4;
	enum { N = 4096 };
	static int32_t xyz[N * 3];
	static uint32_t rgb[N];
	for (int i = 0; i < N; i++) {
		xyz[3 * i] = (i * 7) % 150 - 70; xyz[3 * i + 1] = (i * 13) % 90 - 20; xyz[3 * i + 2] = (i * 29) % 200 - 130;
		rgb[i] = 0x010203u + (uint32_t)i;
	}
	for (int storage = 0; storage < 2; storage++) {
		void* h = create();
		add(h, xyz, rgb, N);
		if (build(h, storage)) return 5;
		uint32_t out[2]; uint8_t ex[2];
		int32_t q[6] = { xyz[3 * (N - 1)], xyz[3 * (N - 1) + 1], xyz[3 * (N - 1) + 2], 1000, 1000, 1000 };
		if (lookup(h, q, 2, out, ex)) return 6;
		if (out[0] != rgb[N - 1] || out[1] != (1u << 30)) { fprintf(stderr, "lookup %x %x\n", out[0], out[1]); return 7; }
	}
	return 0;
}
