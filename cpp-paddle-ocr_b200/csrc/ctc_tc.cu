// tcgen05 CTC head (fc 120->6625 fused with softmax-max / argmax): see launch_ctc_head_tc.
#include "kernels.h"

namespace b200ocr {

bool ctc_tc_eligible(const TV&, int) { return false; }
void launch_ctc_head_tc(const TV&, const __half*, const float*, int, int, int, int*, float*, cudaStream_t) {}

}  // namespace b200ocr
