// How many SMs can a persistent kernel with ~200 KB of shared memory per CTA use at cluster size 1 / 2 / 4 / 8?
// (cudaOccupancyMaxActiveClusters; decides whether 4-CTA clusters -- two UMMA pairs sharing a multicast A tile -- can fill a B200.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(448); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %3d -> %3d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
