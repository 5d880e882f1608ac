// Instantiates the traversal kernel for u8 vectors (both metrics, every lanes-per-row / chunks-per-lane shape).
#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"
namespace fnb {
cudaError_t dispatch_search_u8(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s) {
  return ix->h.metric == FNB_METRIC_IP ? dispatch_gc<DT_U8, M_IP>(ix, p, num_sms, s) : dispatch_gc<DT_U8, M_L2>(ix, p, num_sms, s);
}
}  // namespace fnb
