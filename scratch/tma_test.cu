#include <cstdio>
#include <vector>
#include "../hicpeaks_b200/csrc/hp_device.cuh"
using namespace hp;
__global__ void k(const __grid_constant__ CUtensorMap tm, int q0, int p0, int bytes, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* tile = (double*)smem;
    uint64_t* bar = (uint64_t*)(smem + bytes);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)bytes);
        tma_load_3d(tile, &tm, q0, 0, p0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    double s = 0;
    for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) s += tile[i];
    atomicAdd(out, s);
}
int main() {
    int pitch = 608, num = 71;
    double* d; cudaMalloc(&d, (size_t)pitch * num * 8);
    std::vector<double> h((size_t)pitch * num, 1.0);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    double* out; cudaMalloc(&out, 8);
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    int tests[][2] = {{36, 52}, {38, 52}, {36, 87}, {38, 87}, {40, 87}, {40, 107}, {39, 64}};
    for (auto& t : tests) {
        int nq = t[0], bd = t[1];
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)pitch / 4, 4, (cuuint64_t)num};
        cuuint64_t strides[2] = {(cuuint64_t)pitch / 4 * 8, (cuuint64_t)pitch * 8};
        cuuint32_t box[3] = {(cuuint32_t)nq, 4, (cuuint32_t)bd};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int bytes = nq * 4 * bd * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 64);
        cudaMemset(out, 0, 8);
        k<<<1, 256, bytes + 64>>>(tm, -4, -8, bytes, out);
        cudaError_t e = cudaDeviceSynchronize();
        double ho = 0; cudaMemcpy(&ho, out, 8, cudaMemcpyDeviceToHost);
        printf("nq=%d bd=%d bytes=%d enc=%d -> %s sum=%.0f\n", nq, bd, bytes, (int)r, cudaGetErrorString(e), ho);
        if (e != cudaSuccess) { cudaDeviceReset(); return 1; }
    }
}
