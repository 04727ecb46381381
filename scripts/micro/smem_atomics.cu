// Microbenchmark: cycles per warp-instruction for scattered shared-memory ops on B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters) {
    __shared__ uint32_t tab[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) tab[i] = 0xFFFFFFFFu;
    __syncwarp();
    uint32_t h = (lane * 2654435761u) >> 20;
    long long t0, t1;
    uint32_t acc = 0;
    // 1. dependent LDS chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { acc += ((volatile uint32_t*)tab)[(h + acc) & 4095]; }
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / iters;
    // 2. atomicCAS on scattered addresses (always succeeds: slot empty), dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; i++) { uint32_t o = atomicCAS(&tab[(h + i * 37 + (acc & 1)) & 4095], 0xFFFFFFFFu, 5u + lane); acc += o; }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / iters;
    // 3. atomicCAS that fails (slot taken)
    t0 = clock64();
    for (int i = 0; i < iters; i++) { uint32_t o = atomicCAS(&tab[(h + i * 37 + (acc & 1)) & 4095], 0xFFFFFFFFu, 7u); acc += o; }
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / iters;
    // 4. atomicExch scattered
    t0 = clock64();
    for (int i = 0; i < iters; i++) { uint32_t o = atomicExch(&tab[(h + i * 37 + (acc & 1)) & 4095], 9u); acc += o; }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / iters;
    // 5. ballot + popc dependent
    t0 = clock64();
    for (int i = 0; i < iters; i++) { unsigned b = __ballot_sync(0xffffffffu, (acc + i) & 1); acc += __popc(b); }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / iters;
    // 6. shfl_xor dependent
    t0 = clock64();
    for (int i = 0; i < iters; i++) { acc += __shfl_xor_sync(0xffffffffu, acc, 1); }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / iters;
    // 7. atomicCAS single lane only (lane 0)
    t0 = clock64();
    for (int i = 0; i < iters; i++) { if (lane == 0) { uint32_t o = atomicCAS(&tab[(h + i * 37) & 4095], 0xFFFFFFFFu, 5u); acc += o; } }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / iters;
    // 8. 64-bit LDS dependent
    t0 = clock64();
    { volatile unsigned long long* t64 = (volatile unsigned long long*)tab; unsigned long long a2 = acc;
      for (int i = 0; i < iters; i++) { a2 += t64[(h + a2) & 2047]; } acc += (uint32_t)a2; }
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / iters;
    if (acc == 0x12345) out[15] = acc;
}
int main() {
    long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
    k<<<1, 32>>>(d, 200); cudaDeviceSynchronize();
    k<<<1, 32>>>(d, 200); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    const char* names[] = {"LDS dependent", "ATOMS.CAS success (32 lanes, scattered)", "ATOMS.CAS fail", "ATOMS.EXCH scattered", "ballot+popc", "shfl_xor", "ATOMS.CAS one lane", "LDS.64 dependent"};
    for (int i = 0; i < 8; i++) printf("%-42s %lld cycles\n", names[i], h[i]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
