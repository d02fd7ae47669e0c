#include "common.cuh"
namespace ae {
size_t tc_packed_bytes(int Cs, int Cb, int nsplit) { return (size_t)9 * Cs * Cb * 2 * nsplit; }
int tc_pack_conv(const float*, int, int, int, void*, void*, cudaStream_t) { set_error("tc path not built"); return 1; }
int tc_rowgemm(const RowGemm&, const void*, int, cudaStream_t) { set_error("tc path not built"); return 1; }
int tc_wgrad(const ColGemm&, int, cudaStream_t) { set_error("tc path not built"); return 1; }
bool tc_rowgemm_supported(const RowGemm&) { return false; }
}
