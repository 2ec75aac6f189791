// comm.cu -- multi-GPU execution (one process per GPU).  Placeholder until the NCCL path lands.
#include "common.cuh"

int nfftb_comm_exec_adjoint(nfftb200_plan* p, const void*, void*) { return nfftb_fail(p, NFFTB200_UNSUPPORTED, "node sharding not initialised"); }
int nfftb_comm_exec_forward(nfftb200_plan* p, const void*, void*) { return nfftb_fail(p, NFFTB200_UNSUPPORTED, "node sharding not initialised"); }
void nfftb_comm_destroy(nfftb200_plan*) {}

extern "C" {
int nfftb200_comm_unique_id(void*) { return NFFTB200_UNSUPPORTED; }
int nfftb200_comm_init(nfftb200_plan* p, const void*, int, int, int) { return nfftb_fail(p, NFFTB200_UNSUPPORTED, "NCCL path not built yet"); }
}
