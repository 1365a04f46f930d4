#include "riccati_fast.cuh"

namespace sipoc {
const FastPlan *select_fast_plan(int, int) { return nullptr; }
}  // namespace sipoc
