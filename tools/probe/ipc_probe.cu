// Probe: CUDA IPC peer memory + device-side flag barrier across torchrun ranks (one process per GPU).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return -1; } } while (0)

struct Shared { unsigned long long flags[16]; double vals[2][16]; double data[1024]; };

__global__ void k_allreduce(Shared *mine, Shared *const *peers, int rank, int world, unsigned long long epoch, double partial, double *out, int *status) {
    const int t = threadIdx.x;
    const int slot = (int)(epoch & 1);
    if (t < world) {
        Shared *p = peers[t];
        *((volatile double *)&p->vals[slot][rank]) = partial;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&p->flags[rank]), "l"(epoch) : "memory");
        long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->flags[t]) : "memory");
            if (clock64() - t0 > 4000000000ll) { *status = 1; break; }
        } while (v < epoch);
    }
    __syncthreads();
    if (t == 0) {
        double s = 0;
        for (int i = 0; i < world; i++) s += *((volatile double *)&mine->vals[slot][i]);
        *out = s;
    }
}
__global__ void k_read_peer(Shared *const *peers, int world, double *out) {
    double s = 0;
    for (int i = 0; i < world; i++) s += peers[i]->data[threadIdx.x];
    out[threadIdx.x] = s;
}
__global__ void k_fill(Shared *m, double v) { m->data[threadIdx.x] = v + threadIdx.x; }

static Shared *g_mine; static Shared **g_peers_d; static Shared *g_peers_h[16]; static double *g_out; static int *g_status;
extern "C" int probe_init(int dev, char *handle_out) {
    CK(cudaSetDevice(dev));
    CK(cudaMalloc(&g_mine, sizeof(Shared)));
    CK(cudaMemset(g_mine, 0, sizeof(Shared)));
    CK(cudaMalloc(&g_out, 1024 * sizeof(double)));
    CK(cudaMalloc(&g_status, 4)); CK(cudaMemset(g_status, 0, 4));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, g_mine));
    memcpy(handle_out, &h, sizeof(h));
    CK(cudaDeviceSynchronize());
    return (int)sizeof(h);
}
extern "C" int probe_open(int rank, int world, const char *handles) {
    for (int i = 0; i < world; i++) {
        if (i == rank) { g_peers_h[i] = g_mine; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, handles + i * sizeof(h), sizeof(h));
        CK(cudaIpcOpenMemHandle((void **)&g_peers_h[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    CK(cudaMalloc(&g_peers_d, sizeof(Shared *) * 16));
    CK(cudaMemcpy(g_peers_d, g_peers_h, sizeof(Shared *) * 16, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int probe_run(int rank, int world, int iters, double *sum_out, double *ms_out) {
    k_fill<<<1, 1024>>>(g_mine, 1000.0 * rank);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double s = 0;
    // epoch 1: barrier so that everyone's fill is done
    k_allreduce<<<1, 32>>>(g_mine, g_peers_d, rank, world, 1, (double)(rank + 1), g_out, g_status);
    k_read_peer<<<1, 1024>>>(g_peers_d, world, g_out);
    CK(cudaMemcpy(&s, g_out + 5, 8, cudaMemcpyDeviceToHost));
    sum_out[0] = s;
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) k_allreduce<<<1, 32>>>(g_mine, g_peers_d, rank, world, 2 + i, (double)(rank + 1) * (i + 1), g_out, g_status);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / iters;
    CK(cudaMemcpy(&s, g_out, 8, cudaMemcpyDeviceToHost));
    sum_out[1] = s;
    int st; CK(cudaMemcpy(&st, g_status, 4, cudaMemcpyDeviceToHost));
    return st;
}
