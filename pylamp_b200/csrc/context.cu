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

void plb_ctx_destroy(plb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
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

}  // extern "C"

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
