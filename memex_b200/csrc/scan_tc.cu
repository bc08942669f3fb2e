// scan_tc.cu -- K2 placeholder: the tcgen05 batched scan is not built yet; the store falls back to
// the CUDA-core stream scan for fp16 stores (still a GPU kernel -- there is no CPU path).
#include "common.cuh"
#include "scan.cuh"

namespace mx {
struct TcScanState {};
TcScanState *tc_scan_create(int, uint32_t, uint32_t) { return nullptr; }
void tc_scan_destroy(TcScanState *) {}
void tc_scan_invalidate(TcScanState *) {}
bool tc_scan_supports(const TcScanState *, uint32_t) { return false; }
uint32_t tc_scan_max_k() { return 0; }
uint32_t tc_scan_lists(const TcScanState *, uint64_t) { return 0; }
uint32_t tc_scan_lcap(uint32_t) { return 0; }
cudaError_t tc_scan_launch(TcScanState *, const ScanParams &, uint64_t, uint32_t, KernelTimer *, cudaStream_t,
                           const char **why)
{
    if (why) *why = "tcgen05 scan not built";
    return cudaErrorNotSupported;
}
}  // namespace mx
