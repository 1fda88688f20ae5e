/* ref_stb_shim.c -- the reference's vendored stb_image.h (dependencies/stb/stb_image.h, v2.15) compiled where it lies, behind one door:
 * what glhelper's Texture2D::LoadFromFile calls (texture2d.cpp:37).  TEST INFRASTRUCTURE ONLY: pins host/imageio.cpp. */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_NO_STDIO_WARNINGS
#include "stb_image.h"
#include <string.h>

__attribute__((visibility("default"))) int ref_stbi_load(const char* path, unsigned char* rgba, int capacity_pixels, int* w, int* h, int* comps)
{
	int x = 0, y = 0, n = 0;
	stbi_uc* data = stbi_load(path, &x, &y, &n, 4);
	if (!data) return 1;
	*w = x; *h = y; *comps = n;
	if (capacity_pixels >= x * y && rgba) memcpy(rgba, data, (size_t)x * y * 4);
	stbi_image_free(data);
	return 0;
}
