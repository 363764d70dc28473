// include/corto/decoder.h — forwarding header: user code written against the reference (crt::Decoder) compiles unchanged with
// -I<corto-b200>/include; everything lives in corto_b200/decoder.h (decode side only, CUDA behind the C ABI).
#include "../corto_b200/decoder.h"
