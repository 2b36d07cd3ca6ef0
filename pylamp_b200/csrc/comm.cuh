// comm.cuh -- slab communicator helpers (see comm.cu)
#pragma once
#include "common.cuh"

enum { PLB_OP_SUM = 0, PLB_OP_MAX = 1, PLB_OP_MIN = 2 };

int plb_comm_rank(const plb_ctx* ctx);
int plb_comm_size(const plb_ctx* ctx);
// in-place all-reduce of `count` doubles on the context's stream (no-op for a single rank)
int plb_comm_allreduce(plb_ctx* ctx, double* d_buf, size_t count, int op);
// one halo row per plane with each z-neighbour (no-op for a single rank)
int plb_comm_halo_exchange(plb_ctx* ctx, double* base, int nplanes, size_t plane_stride, int ld, int ncols,
                           int lo, int r0, int r1);
