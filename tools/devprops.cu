#include <cstdio>
#include <cuda_runtime.h>
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("name %s sm %d.%d SMs %d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    printf("l2CacheSize %d persistingL2CacheMaxSize %d accessPolicyMaxWindowSize %d\n", p.l2CacheSize, p.persistingL2CacheMaxSize, p.accessPolicyMaxWindowSize);
    printf("sharedMemPerMultiprocessor %zu regsPerMultiprocessor %d maxThreadsPerMultiProcessor %d\n", p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor, p.maxThreadsPerMultiProcessor);
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize); printf("default persisting limit %zu\n", lim);
    printf("totalGlobalMem %zu memoryBusWidth %d memClock %d\n", p.totalGlobalMem, p.memoryBusWidth, p.memoryClockRate);
    return 0;
}
