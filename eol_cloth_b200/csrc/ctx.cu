// ctx.cu — context + error plumbing + host-side mesh flattening helper of the C ABI (include/eolc.h).
#include "common.h"
#include <unordered_map>

namespace eolc {
static thread_local std::string g_last_error;
void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
}  // namespace eolc

extern "C" {

const char *eolc_last_error(void) { return eolc::g_last_error.c_str(); }

int eolc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int eolc_ctx_create(int device, eolc_ctx **out) {
    EOLC_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        eolc::set_error("no CUDA device available (%s); this library has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return EOLC_ERR_CUDA;
    }
    EOLC_REQUIRE(device >= 0 && device < n, "device index out of range");
    EOLC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    EOLC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        eolc::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return EOLC_ERR_UNSUPPORTED;
    }
    eolc_ctx *c = new eolc_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t es = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) {
        delete c;
        eolc::set_error("cudaStreamCreateWithFlags failed: %s", cudaGetErrorString(es));
        return EOLC_ERR_CUDA;
    }
    // optional: without them the host entries copy on the one stream
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); c->copy_stream = nullptr; }
    if (cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); c->copy_event = nullptr; }
    *out = c;
    return EOLC_OK;
}

void *eolc_host_alloc(size_t bytes) {
    if (bytes == 0) return nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        eolc::set_error("no CUDA device available (%s); this library has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    void *p = nullptr;
    // portable: page-locked for every CUDA context of the process (one ctx per GPU in an ensemble host)
    e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); eolc::set_error("eolc_host_alloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

void eolc_host_free(void *p) { if (p) cudaFreeHost(p); }

// A caller's OWN array (e.g. the value array of an Eigen::SparseMatrix) becomes page-locked in place, so that the host entry points
// DMA straight into it; undone by eolc_host_unregister before the array is freed or reallocated.
int eolc_host_register(void *p, size_t bytes) {
    EOLC_REQUIRE(p && bytes, "eolc_host_register: NULL / empty range");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return EOLC_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); eolc::set_error("eolc_host_register(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); return EOLC_ERR_CUDA; }
    return EOLC_OK;
}
int eolc_host_unregister(void *p) {
    if (!p) return EOLC_OK;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); if (e != cudaErrorHostMemoryNotRegistered) { eolc::set_error("eolc_host_unregister failed: %s", cudaGetErrorString(e)); return EOLC_ERR_CUDA; } }
    return EOLC_OK;
}

void eolc_ctx_destroy(eolc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
    delete ctx;
}

void *eolc_ctx_stream(eolc_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// Mesh::add(Face) edge creation, /root/reference/src/external/ArcSim/mesh.cpp:356-378:
// for face k, i = 0,1,2: edge (v[i] -> v[NEXT(i)]) is appended iff no edge joins those nodes yet, with
// n[0] = v[i], n[1] = v[NEXT(i)]; adjf[side], side = 0 iff the face runs n[0] -> n[1].
// Stencil = (n0, n1, opp(adjf[0]), opp(adjf[1]))  (get_other_vert, mesh.hpp:276-280; Forces.cpp:692-697).
int eolc_mesh_edge_stencils(int32_t N, int32_t F, const int32_t *face_nodes, int32_t *E_out, int32_t *edge_stencil) {
    EOLC_REQUIRE(N >= 0 && F >= 0 && face_nodes && E_out && edge_stencil, "bad arguments");
    std::unordered_map<uint64_t, int32_t> edge_of;
    edge_of.reserve((size_t)F * 2 + 16);
    int32_t E = 0;
    for (int32_t k = 0; k < F; ++k) {
        const int32_t *v = face_nodes + 3 * (size_t)k;
        for (int i = 0; i < 3; ++i) EOLC_REQUIRE(v[i] >= 0 && v[i] < N, "face node index out of range");
        for (int i = 0; i < 3; ++i) {  // add_edges_if_needed
            int32_t a = v[i], b = v[(i + 1) % 3];
            uint64_t key = ((uint64_t)(uint32_t)(a < b ? a : b) << 32) | (uint32_t)(a < b ? b : a);
            if (edge_of.find(key) == edge_of.end()) {
                edge_of[key] = E;
                edge_stencil[4 * (size_t)E + 0] = a;
                edge_stencil[4 * (size_t)E + 1] = b;
                edge_stencil[4 * (size_t)E + 2] = -1;
                edge_stencil[4 * (size_t)E + 3] = -1;
                ++E;
            }
        }
        for (int i = 0; i < 3; ++i) {  // adjacency: edge opposite v[i] joins v[NEXT(i)] -> v[PREV(i)]
            int32_t v0 = v[(i + 1) % 3], v1 = v[(i + 2) % 3];
            uint64_t key = ((uint64_t)(uint32_t)(v0 < v1 ? v0 : v1) << 32) | (uint32_t)(v0 < v1 ? v1 : v0);
            int32_t e = edge_of[key];
            int side = edge_stencil[4 * (size_t)e] == v0 ? 0 : 1;
            edge_stencil[4 * (size_t)e + 2 + side] = v[i];  // the face's other vertex
        }
    }
    *E_out = E;
    return EOLC_OK;
}

}  // extern "C"
