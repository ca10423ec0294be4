// TEST INFRASTRUCTURE ONLY - never loaded by the transcriptclean_b200 package.
// CPU build of the engine's C ABI: the planning code (csrc/tc_plan.cuh) is shared verbatim,
// kernels become serial loops and the warp-level output stage is replaced by a scalar twin.
// It lets the host-side parser / renderer / CLI and the planning semantics be tested against
// the oracle on machines without a GPU.  GPU parity is established separately (-m gpu tests).
#define TC_HOSTSIM 1
#include "../../transcriptclean_b200/csrc/tc_b200.cu"
#include "../../transcriptclean_b200/csrc/tc_host.cpp"
#include "../../transcriptclean_b200/csrc/tc_tables.cpp"
