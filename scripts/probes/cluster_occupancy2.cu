// cudaOccupancyMaxActiveClusters as a function of shared memory per CTA (i.e. CTAs per SM), block size and cluster size: do
// clustered launches get more than one CTA per SM co-resident?  (decides whether 256 CTAs in clusters of 8 can all be running
// at once)   usage: cluster_occupancy2 [threads] [smem bytes ...]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
__global__ void __launch_bounds__(384) k384(int* p) { extern __shared__ __align__(1024) char s2[]; if (p) p[0] = s2[0]; }
template <typename K> void probe(K* kern, const char* name, int threads, int bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int cs : {1, 2, 4, 8}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1184 / cs * cs); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = bytes;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        printf("%s threads %d smem %6d B  cluster size %2d: max active clusters %3d -> %3d CTAs (%s)\n", name, threads, bytes, cs, n, n * cs, cudaGetErrorString(e));
    }
}
int main(int argc, char** argv) {
    for (int bytes : {204800, 116736, 115712, 115056, 114688, 106864}) {
        probe(k, "plain  ", 256, bytes);
        probe(k, "plain  ", 384, bytes);
        probe(k384, "aligned", 384, bytes);
    }
    return 0;
}
