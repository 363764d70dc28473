// tests/cpp/custom_attr_main.cpp — setAttribute(name, buffer, VertexAttribute*) (src/decoder.cpp:104-114): an object of the
// stream's own codec is adopted (ownership moves to the decoder), a user codec is refused with a thrown message.
#include <stdio.h>
#include <string.h>
#include <vector>
#include <corto/decoder.h>

struct MyCodec: public crt::VertexAttribute { virtual int codec() { return CUSTOM_CODEC; } };

int main(int argc, char **argv) {
	if(argc < 2) return 2;
	FILE *f = fopen(argv[1], "rb");
	if(!f) return 2;
	fseek(f, 0, SEEK_END); long len = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<uint32_t> store((len + 3)/4 + 1);
	if(fread(store.data(), 1, len, f) != (size_t)len) return 2;
	fclose(f);
	crt::Decoder decoder((int)len, (const unsigned char *)store.data());
	std::vector<float> n(decoder.nvert*3);
	crt::NormalAttr *mine = new crt::NormalAttr();
	if(!decoder.setAttribute("normal", (char *)n.data(), mine)) return 3;          // same codec: adopted
	if(decoder.data["normal"] != mine) return 4;
	if(decoder.setAttribute("nosuch", (char *)n.data(), (crt::VertexAttribute *)nullptr)) return 5;   // absent: false, like the reference
	try {
		decoder.setAttribute("normal", (char *)n.data(), new MyCodec());
	} catch(const char *msg) {
		printf("refused: %s\n", msg);
		return strstr(msg, "custom attribute codecs") ? 0 : 6;
	}
	return 7;
}
