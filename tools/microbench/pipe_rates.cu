// Dev tool: issue cost (clocks per warp instruction and SM sub-partition) of the integer instructions the packed alignment
// kernel is made of, on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, CH = 8;

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c, uint32_t& acc)
{
    if (OP == 0) return __vmaxs2(a, b);                  // VIMNMX.S16x2
    if (OP == 1) return __viaddmax_s16x2(a, b, c);       // VIADDMNMX.S16x2
    if (OP == 2) return __vimax3_s16x2(a, b, c);         // VIMNMX3.S16x2
    if (OP == 3) return a * 3u + b;                      // IMAD
    if (OP == 4) return (a + b) ^ c;                     // IADD + LOP3 (2 instr)
    if (OP == 5) {                                       // VIMNMX.S16x2 with predicate outputs + 2 predicated IMAD
	uint32_t val;
	asm("{.reg .pred pu, pv; .reg .s16 rs0, rs1, rs2, rs3; max.s16x2 %0, %2, %3; mov.b32 {rs0, rs1}, %0; mov.b32 {rs2, rs3}, %2;"
	    "setp.eq.s16 pv, rs0, rs2; setp.eq.s16 pu, rs1, rs3; @!pv mad.lo.u32 %1, %4, 16, %1; @!pu mad.lo.u32 %1, %4, 4096, %1;}"
	    : "=r"(val), "+r"(acc) : "r"(a), "r"(b), "r"(c));
	return val;
    }
    if (OP == 6) return max(int(a), int(b));             // IMNMX (32-bit)
    if (OP == 7) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }  // PRMT
    if (OP == 8) return __vmins2(a - b, 0x00010001u);    // sub + VIMNMX
    if (OP == 9) return __viaddmax_s32(a, b, c);         // VIADDMNMX 32-bit
    return a;
}

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* clk)
{
    uint32_t v[CH], acc = 0;
    for (int i = 0; i < CH; i++) v[i] = seed * (threadIdx.x + 1) + i * 77u;
    const uint32_t b = seed | 0x00030003u, c = seed | 1u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
	for (int i = 0; i < CH; i++) v[i] = op<OP>(v[i], b, c, acc);
    }
    const long long t1 = clock64();
    uint32_t s = acc;
    for (int i = 0; i < CH; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_op)
{
    uint32_t* out;
    long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 8);
    for (int warps_per_smsp : {1, 2, 4, 8}) {
	const int threads = 32 * 4 * warps_per_smsp;
	k<OP><<<148, threads>>>(out, 12345u, clk);
	k<OP><<<148, threads>>>(out, 12345u, clk);
	cudaDeviceSynchronize();
	long long h = 0;
	cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
	// clocks per warp instruction per sub-partition = elapsed / (instructions issued per sub-partition)
	const double per = double(h) / (double(ITER) * CH * instr_per_op * warps_per_smsp);
	printf("%-44s warps/SMSP %d: %7.2f clk per warp-instr (per op %.2f)\n", name, warps_per_smsp, per, per * instr_per_op);
    }
    cudaFree(out);
    cudaFree(clk);
}

int main()
{
    run<0>("VIMNMX.S16x2", 1);
    run<1>("VIADDMNMX.S16x2", 1);
    run<2>("VIMNMX3.S16x2", 1);
    run<3>("IMAD", 1);
    run<4>("IADD3+LOP3 (2 instr)", 2);
    run<5>("VIMNMX.S16x2 +2 preds +2 pred IMAD (3 instr)", 3);
    run<6>("IMNMX s32", 1);
    run<7>("PRMT", 1);
    run<8>("sub + VIMNMX.S16x2 (2 instr)", 2);
    run<9>("VIADDMNMX s32", 1);
    return 0;
}
