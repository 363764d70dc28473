// tests/cpp/facade_main.cpp — a user program written against the reference's README usage (README.md:212-228), compiled
// against include/corto_b200/decoder.h instead of <corto/decoder.h>.  Decodes argv[1] and dumps the arrays to argv[2].
#include <stdio.h>
#include <vector>
#include <corto_b200/decoder.h>

int main(int argc, char **argv) {
	if(argc < 3) return 2;
	FILE *f = fopen(argv[1], "rb");
	if(!f) return 2;
	fseek(f, 0, SEEK_END); long len = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<uint32_t> store((len + 3)/4 + 1);                 // 4-byte aligned, as the reference demands
	if(fread(store.data(), 1, len, f) != (size_t)len) return 2;
	fclose(f);
	try {
		crt::Decoder decoder((int)len, (const uchar *)store.data());
		std::vector<float> coords(decoder.nvert*3), normals(decoder.nvert*3), uvs(decoder.nvert*2);
		std::vector<uchar> colors(decoder.nvert*4);
		std::vector<uint32_t> index(decoder.nface*3);
		decoder.setPositions(coords.data());
		bool hn = decoder.setNormals(normals.data());
		bool hc = decoder.setColors(colors.data(), 4);
		bool hu = decoder.setUvs(uvs.data());
		if(decoder.nface) decoder.setIndex(index.data());
		decoder.decode();
		FILE *o = fopen(argv[2], "wb");
		uint32_t hdr[6] = { decoder.nvert, decoder.nface, hn, hc, hu, (uint32_t)decoder.index.groups.size() };
		fwrite(hdr, 4, 6, o);
		fwrite(coords.data(), 4, coords.size(), o);
		if(hn) fwrite(normals.data(), 4, normals.size(), o);
		if(hc) fwrite(colors.data(), 1, colors.size(), o);
		if(hu) fwrite(uvs.data(), 4, uvs.size(), o);
		fwrite(index.data(), 4, index.size(), o);
		fclose(o);
	} catch(const char *error) {
		fprintf(stderr, "corto: %s\n", error);
		return 1;
	}
	return 0;
}
