// placeholder — replaced by the tcgen05 kernel
#include "bkm_common.cuh"
namespace bkm {
bool tc_supported(int, int, int) { return false; }
int launch_tc(const ChunkArgs&, bool, int, int*, cudaStream_t) { return BKM_EUNSUPPORTED; }
}
