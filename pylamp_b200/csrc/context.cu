// context.cu -- plb_ctx lifetime, stream binding, scratch memory.
#include "common.cuh"

extern "C" {

const char* plb_version(void) { return "pylamp_b200 0.1.0 (sm_100a, fp64)"; }

int plb_ctx_create(int device, plb_ctx** out) {
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return 3;
    if (cudaSetDevice(device) != cudaSuccess) return 3;
    plb_ctx* c = new plb_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return 3;
    }
    c->own_stream = true;
    c->t2g_variant = 1;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    if (cudaMallocHost((void**)&c->h_pinned, 64 * sizeof(double)) != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return 3;
    }
    *out = c;
    return 0;
}

extern "C" void plb_comm_destroy(plb_ctx* ctx);
extern "C" void plb_migrate_free(plb_ctx* ctx);
extern "C" void plb_inject_free(plb_ctx* ctx);

void plb_ctx_destroy(plb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    plb_migrate_free(ctx);
    plb_inject_free(ctx);
    plb_comm_destroy(ctx);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->slab_scratch) cudaFree(ctx->slab_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * ctx->prof_cap; i++) cudaEventDestroy(ctx->prof_ev[i]);
        delete[] ctx->prof_ev;
        delete[] ctx->prof_cls;
        delete[] ctx->prof_bytes;
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int plb_ctx_set_stream(plb_ctx* ctx, void* cuda_stream) {
    // the handle is used as given; 0 is CUDA's default stream (what torch uses unless told otherwise)
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) {
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int plb_ctx_sync(plb_ctx* ctx) {
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

const char* plb_last_error(plb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

long long plb_launch_count(plb_ctx* ctx) { return ctx ? ctx->launches : -1; }

int plb_ctx_set_param(plb_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return 1;
    if (!strcmp(name, "t2g_variant")) {
        ctx->t2g_variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "t2g_parts")) {
        ctx->t2g_parts = (int)value;
        return 0;
    }
    if (!strcmp(name, "t2g_nm")) {
        ctx->t2g_nm = (int)value;
        return 0;
    }
    if (!strcmp(name, "t2g_minb")) {
        ctx->t2g_minb = (int)value;
        return 0;
    }
    if (!strcmp(name, "t2g_nfmax")) {
        ctx->t2g_nfmax = (int)value;
        return 0;
    }
    PLB_FAIL(ctx, "plb_ctx_set_param: unknown parameter '%s'", name);
}

int plb_profile_enable(plb_ctx* ctx, int on) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on && !ctx->prof_ev) {
        ctx->prof_cap = 16384;
        ctx->prof_ev = new cudaEvent_t[2 * ctx->prof_cap];
        ctx->prof_cls = new int[ctx->prof_cap];
        ctx->prof_bytes = new double[ctx->prof_cap];
        for (int i = 0; i < 2 * ctx->prof_cap; i++) PLB_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[i]));
    }
    ctx->prof_used = 0;
    for (int i = 0; i < 16; i++) ctx->prof_skipped[i] = 0;
    ctx->prof_on = on != 0;
    return 0;
}

int plb_profile_read(plb_ctx* ctx, long long* h_count, double* h_ms, double* h_bytes) {
    if (!ctx) return 1;
    PLB_CUDA(ctx, cudaSetDevice(ctx->device));
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < PLB_K_NCLASS; i++) h_count[i] = 0, h_ms[i] = 0, h_bytes[i] = 0;
    for (int i = 0; i < ctx->prof_used; i++) {
        float ms = 0;
        PLB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        h_count[ctx->prof_cls[i]]++;
        h_ms[ctx->prof_cls[i]] += ms;
        h_bytes[ctx->prof_cls[i]] += ctx->prof_bytes[i];
    }
    ctx->prof_used = 0;
    return 0;
}

}  // extern "C"

void plb_prof_begin(plb_ctx* ctx, int cls, double bytes) {
    if (ctx->prof_used >= ctx->prof_cap) {
        ctx->prof_skipped[cls]++;
        return;
    }
    ctx->prof_cls[ctx->prof_used] = cls;
    ctx->prof_bytes[ctx->prof_used] = bytes;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used], ctx->stream);
}

void plb_prof_end(plb_ctx* ctx) {
    if (ctx->prof_used >= ctx->prof_cap) return;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used + 1], ctx->stream);
    ctx->prof_used++;
}

int plb_ws_reserve(plb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return 0;
    PLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ws) PLB_CUDA(ctx, cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = bytes + bytes / 4;
    PLB_CUDA(ctx, cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
    return 0;
}
