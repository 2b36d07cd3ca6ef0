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

// ---- slab-local fields: full-size (nz x ld) arrays of which a rank keeps only its own node rows
// [ctx->slab_i0, ctx->slab_i1) plus `ctx->slab_halo` rows of each neighbour current.
// Exchange `h` halo rows of `narr` full-size arrays (row length row_doubles[a]) with both z-neighbours:
// my first/last h own rows go out, the neighbours' rows land in [i0-h, i0) and [i1, i1+h).  One NCCL group.
int plb_comm_halo_rows(plb_ctx* ctx, int narr, double* const* arrs, const long long* row_doubles, int i0, int i1, int h);
// Boundary-row accumulate (marker->node sums of slab-owned markers): the rows [i0-h, i0) and [i1, i1+h) of
// every array, which hold this rank's contributions to its neighbours' nodes, are added to the neighbours'
// rows and theirs to mine.  `scratch`: 2*h*sum(row_doubles) doubles of device memory.
int plb_comm_accumulate_rows(plb_ctx* ctx, int narr, double* const* arrs, const long long* row_doubles,
                             const int* i0, const int* i1, const int* nrows, int h, double* scratch);
// Exchange with the two z-neighbours in one NCCL group: send_dn -> rank-1, recv_dn <- rank-1, send_up -> rank+1,
// recv_up <- rank+1 (counts in doubles; a count of 0 skips that transfer -- the peer's matching count is 0 too).
int plb_comm_neighbour_exchange(plb_ctx* ctx, const double* send_dn, size_t n_send_dn, double* recv_dn, size_t n_recv_dn,
                                const double* send_up, size_t n_send_up, double* recv_up, size_t n_recv_up);
