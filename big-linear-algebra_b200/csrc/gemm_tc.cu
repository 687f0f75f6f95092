// gemm_tc.cu -- placeholder until the tcgen05 3xTF32 kernel lands (next commit).
#include "kernels.h"
namespace bla {
bool gemm_3xtf32(const GemmArgs&, cudaStream_t) { return false; }
}
