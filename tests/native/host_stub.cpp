// tests/native/host_stub.cpp -- TEST ONLY.  Lets the host side of the product (centrifuge_b200/csrc/cf_host.cpp +
// cf_index.cpp) link without the CUDA translation units, so that its CPU-testable entry points (cfb_test_parse,
// cfb_test_host_path, cfb_kreport, cfb_em_abundance_host) can be run under ASan / UBSan:
//   g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=address,undefined -shared -fPIC -o /tmp/libcfbhost_san.so \
//       tests/native/host_stub.cpp centrifuge_b200/csrc/cf_host.cpp centrifuge_b200/csrc/cf_index.cpp -lpthread
//   LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
//       CFB_PRODUCT_LIB=/tmp/libcfbhost_san.so python -m pytest tests/test_reader_fuzz.py tests/test_host_path.py tests/test_kreport.py tests/test_em_host.py -m "not gpu"
// Every device entry point fails with CFB_ENODEV here: this is not a CPU fallback and nothing in the product uses it.
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/cfb200.h"
#include "../../centrifuge_b200/csrc/cf_index.h"

struct cfb_index { cfb::HostIndex h; };
struct cfb_ctx { int unused; };
static std::string g_err;

extern "C" {
const char* cfb_last_error(void) { return g_err.c_str(); }
int cfb_index_load_ex(const char* base, int device, uint32_t, cfb_index** out) {
	if(device >= 0) { g_err = "host stub: no device"; return CFB_ENODEV; }
	cfb_index* ix = new cfb_index();
	const std::string e = cfb::load_cf_index(base, ix->h);
	if(!e.empty()) { g_err = e; delete ix; return CFB_EIO; }
	*out = ix; return CFB_OK;
}
int cfb_index_load(const char* base, int device, cfb_index** out) { return cfb_index_load_ex(base, device, 0u, out); }
void cfb_index_free(cfb_index* ix) { delete ix; }
const cfb::HostIndex* cfb_index_host(const cfb_index* ix) { return ix ? &ix->h : NULL; }
void cfb_params_default(cfb_params* p) { memset(p, 0, sizeof *p); p->khits = 5; p->min_hitlen = 22; p->tree_traverse = 1; }
int cfb_ctx_create(const cfb_index*, const cfb_params*, cfb_ctx**) { g_err = "host stub: no device"; return CFB_ENODEV; }
void cfb_ctx_destroy(cfb_ctx*) {}
int cfb_ctx_slots(const cfb_ctx*) { return 0; }
int cfb_classify_submit(cfb_ctx*, int, const cfb_batch*) { return CFB_ENODEV; }
int cfb_classify_wait(cfb_ctx*, int, cfb_result*) { return CFB_ENODEV; }
int cfb_text_submit(cfb_ctx*, int, const void*, uint64_t, const void*, uint64_t, uint64_t, const cfb_text_opts*) { return CFB_ENODEV; }
int cfb_text_wait(cfb_ctx*, int, int, cfb_text_result*) { return CFB_ENODEV; }
int cfb_text_species(cfb_ctx*, uint64_t*, uint64_t*, uint64_t*, uint64_t*, uint64_t, uint64_t* n) { if(n) *n = 0; return CFB_ENODEV; }
int cfb_em_abundance(int, uint64_t, uint64_t, const uint64_t*, const uint64_t*, const uint32_t*, const uint64_t*, double*, uint64_t*, double*) { return CFB_ENODEV; }
const char* cfb_em_last_error(void) { return "host stub: no device"; }
int cfb_counts_read(cfb_ctx*, int, uint64_t*, uint64_t*, uint64_t*, uint64_t*, uint64_t, uint64_t* n) { if(n) *n = 0; return CFB_ENODEV; }
int cfb_counts_allreduce(cfb_ctx* const*, int, uint64_t*, uint64_t) { return CFB_ENODEV; }
int cfb_comm_init_all(cfb_ctx* const*, int) { return CFB_ENODEV; }
int cfb_device_count(void) { return 0; }
void* cfb_host_alloc(size_t n) { return malloc(n ? n : 1); }
void cfb_host_free(void* p) { free(p); }
}
