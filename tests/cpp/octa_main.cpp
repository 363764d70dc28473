// tests/cpp/octa_main.cpp — user code calling the NormalAttr statics (normal_attribute.h:75-122) through the REFERENCE's include
// name.  Built twice by tests/test_facade.py: with -I<this repo>/include/corto (forwarding headers) and, where /root/reference
// exists, with the reference's own include directory; the two binaries must print the same bytes.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <normal_attribute.h>

static uint32_t rng(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

int main() {
	uint32_t s = 12345u, h = 2166136261u;
	auto mix = [&](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; for(size_t i = 0; i < n; i++) { h ^= b[i]; h *= 16777619u; } };
	const int units[3] = {511, 127, 32767};
	for(int it = 0; it < 200000; it++) {
		float f[3];
		for(int k = 0; k < 3; k++) f[k] = ((int)(rng(s) % 2001) - 1000)/(float)(1 + rng(s) % 7);
		if(it % 97 == 0) f[2] = 0.f;
		if(it % 1013 == 0) { f[0] = f[1] = 0.f; }
		const int unit = units[it % 3];
		if(f[0] != 0.f || f[1] != 0.f || f[2] != 0.f) {            // (0,0,0) is 0/0 = NaN -> (int)NaN: x86 gives INT_MIN, but it is UB in C++
			crt::Point2i o = crt::NormalAttr::toOcta(crt::Point3f(f[0], f[1], f[2]), unit);
			int32_t oo[2] = {o[0], o[1]}; mix(oo, 8);
		}
		crt::Point3i vi((int)(rng(s) % 401) - 200, (int)(rng(s) % 401) - 200, (int)(rng(s) % 401) - 200);
		crt::Point2i oi = crt::NormalAttr::toOcta(vi, unit);
		int32_t oo2[2] = {oi[0], oi[1]}; mix(oo2, 8);
		const int a = (int)(rng(s) % (2u*unit + 1u)) - unit, b = (int)(rng(s) % (2u*unit + 1u)) - unit;
		crt::Point3f n = crt::NormalAttr::toSphere(crt::Point2i(a, b), unit);
		float nn[3] = {n[0], n[1], n[2]}; mix(nn, 12);
		crt::Point3s q = crt::NormalAttr::toSphere(crt::Point2s((int16_t)a, (int16_t)b), unit);
		int16_t qq[3] = {q[0], q[1], q[2]}; mix(qq, 6);
	}
	printf("%08x\n", h);
	return 0;
}
