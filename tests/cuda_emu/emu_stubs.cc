// tests/cuda_emu/emu_stubs.cc -- TEST INFRASTRUCTURE: symbols the emulated subset expects from product files that are not part of it
#include "mmb_internal.h"
extern "C" mmb_ctx_t *mmb_ctx_create(int device);
mmb_ctx_t *mmb_default_ctx(void) { static mmb_ctx_t *c = mmb_ctx_create(0); return c; }
