// v1 placeholder: the TMA-staged thread-per-column kernel registers here.
#include "b2b_internal.h"
int b2b_launch_chain_v1(const B2BChainParams&, cudaStream_t) { return B2B_EUNSUPPORTED; }
int b2b_chain_grid_size_v1(const B2BChainParams&) { return 0; }
