// temporary stub: replaced by hnsw_host.cpp + hnsw_search.cu
#include "hnsw.h"
namespace vl {
struct HnswState { int dummy; };
void HnswDeleter::operator()(HnswState* s) const { delete s; }
HnswState* hnsw_state_create(uint32_t, int, uint32_t, uint32_t, uint32_t) { return nullptr; }
void hnsw_state_release_device(HnswState*) {}
bool hnsw_has_id(const HnswState*, uint64_t) { return false; }
bool hnsw_index_of(const HnswState*, uint64_t, uint64_t*) { return false; }
bool hnsw_max_id(const HnswState*, uint64_t*) { return false; }
uint64_t hnsw_live(const HnswState*) { return 0; }
int hnsw_add_rows(HnswState*, const uint64_t*, const float*, uint64_t) { return 9; }
bool hnsw_soft_delete(HnswState*, uint64_t) { return false; }
int hnsw_upload(HnswState*, cudaStream_t) { return 9; }
int hnsw_search_host(HnswState*, const float*, uint32_t, const float*, uint32_t, uint32_t, uint32_t, uint64_t*, double*, uint32_t*, cudaStream_t, uint64_t*, uint64_t*) { return 9; }
}
