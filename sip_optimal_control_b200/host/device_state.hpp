// Private to the host shim: what hangs off LQR::Workspace::device.
#pragma once

#include <vector>

#include "../../include/sipoc.h"
#include "lqr.hpp"

namespace sip::optimal_control {

struct LQR::DeviceState {
  sipoc_engine *engine = nullptr;
  std::vector<double> in[9];   // Q M R q r A B c delta, flat
  std::vector<double> out[3];  // x u y, flat
  sipoc_error last_error = SIPOC_OK;  // latched: the reference's void methods cannot return it
  ~DeviceState() {
    if (engine != nullptr) sipoc_destroy(engine);
  }
};

}  // namespace sip::optimal_control
