// flat_tensor.cu -- tensor-core candidate pass (placeholder until the tcgen05 kernel lands).
#include "flat_index.cuh"

namespace cm {

bool FlatIndex::tensor_path_eligible(int64_t, int64_t, bool, float) const { return false; }

int FlatIndex::search_tensor(const float *, int64_t, int64_t, const uint8_t *, float, int64_t, uint32_t *, float *,
                             int64_t *, int64_t *, cudaStream_t, cm_flat_stats *) {
    return fail(CM_ERR_UNSUPPORTED, "tensor path not built yet");
}

void FlatIndex::free_shadow() {
    cudaFree(rows_bf16);
    cudaFree(row_sqnorm);
    rows_bf16 = nullptr; row_sqnorm = nullptr; shadow_rows = shadow_cap = 0;
}

}  // namespace cm
