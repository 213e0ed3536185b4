// Dev probe (GPU box, under ncu): how does sm_100a count wavefronts for a warp-wide STS.64 / predicated STS.64?
// One launch per pattern; ncu reports l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum per launch, the host prints
// the half-warp model's prediction (sum over the two half-warps of the largest number of distinct 8-byte words that
// share a bank pair).   nvcc -arch=sm_100a -o sts_bank_probe sts_bank_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <set>
#include <map>
__global__ void probe(const int *__restrict__ word, const int *__restrict__ on, double *out, int reps)
{
    __shared__ double xs[4096];
    const int l = threadIdx.x;
    for (int i = l; i < 4096; i += 32) xs[i] = 0.0;
    __syncwarp();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(xs);
    const uint32_t a = base + 8u * (uint32_t)word[l];
    const uint32_t p = (uint32_t)on[l];
    for (int r = 0; r < reps; ++r)
        asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f64 [%0], %1;}" ::"r"(a), "d"((double)(l + r)), "r"(p) : "memory");
    __syncwarp();
    out[l] = xs[word[l]];
}
static int model(const std::vector<int> &w, const std::vector<int> &on)
{
    int tot = 0;
    for (int h = 0; h < 2; ++h) {
        std::map<int, std::set<int>> by;
        for (int l = 16 * h; l < 16 * h + 16; ++l) if (on[l]) by[w[l] & 15].insert(w[l]);
        int m = 0; for (auto &kv : by) m = std::max<int>(m, (int)kv.second.size());
        tot += m;
    }
    return tot;
}
int main()
{
    int *d_w, *d_on; double *d_out;
    cudaMalloc(&d_w, 128); cudaMalloc(&d_on, 128); cudaMalloc(&d_out, 256);
    std::vector<std::vector<int>> W, ON; std::vector<const char *> name;
    auto add = [&](const char *n, std::vector<int> w, std::vector<int> on = std::vector<int>(32, 1)) { W.push_back(w); ON.push_back(on); name.push_back(n); };
    std::vector<int> w(32), on(32, 1);
    for (int l = 0; l < 32; ++l) w[l] = l;              add("0 identity", w);
    for (int l = 0; l < 32; ++l) w[l] = l & 15;         add("1 halves write the same 16 words", w);
    for (int l = 0; l < 32; ++l) w[l] = 0;              add("2 all one word", w);
    for (int l = 0; l < 32; ++l) w[l] = (l & 1) ? 16 + l / 2 : l / 2;   add("3 2-way conflict inside each half (w, w+16)", w);
    for (int l = 0; l < 32; ++l) w[l] = 16 * (l & 15);  add("4 16-way conflict per half", w);
    for (int l = 0; l < 32; ++l) w[l] = (l < 16) ? l : 16 * (l - 16) ;  add("5 half 0 clean, half 1 16-way", w);
    for (int l = 0; l < 32; ++l) w[l] = (l < 8) ? l : (l < 16 ? 16 + (l - 8) : l + 16);  add("6 half 0: words 0-7 and 16-23 (2-way), half 1 clean", w);
    for (int l = 0; l < 32; ++l) w[l] = 18 * l;         add("7 stride 18", w);
    for (int l = 0; l < 32; ++l) w[l] = l; for (int l = 0; l < 32; ++l) on[l] = l < 16; add("8 identity, half 1 predicated off", w, on);
    for (int l = 0; l < 32; ++l) on[l] = (l & 1);       add("9 identity, even lanes off", w, on);
    for (int l = 0; l < 32; ++l) on[l] = (l % 4 == 0);  add("10 identity, every 4th lane on", w, on);
    for (int l = 0; l < 32; ++l) on[l] = l == 5;        add("11 identity, one lane on", w, on);
    for (int l = 0; l < 32; ++l) { w[l] = (l < 16) ? l : (l & 7); on[l] = 1; } add("12 half 1: lanes pairwise same word (8 words twice)", w, on);
    for (int l = 0; l < 32; ++l) { w[l] = (l & 7); } add("13 every word four times", w, on);
    for (int l = 0; l < 32; ++l) { w[l] = (l < 16) ? 2 * l : 2 * (l - 16) + 1; } add("14 half 0 even words 0..30 (pairs w,w+16 collide), half 1 odd", w, on);
    for (int l = 0; l < 32; ++l) { w[l] = ((l >> 1) & 7) + 16 * (l & 1) + 32 * (l >> 4); } add("15 neighbours lanes collide (w, w+16)", w, on);
    for (int l = 0; l < 32; ++l) { w[l] = (l & 7) + 16 * ((l >> 3) & 1) + 32 * (l >> 4); } add("16 lanes l, l+8 collide", w, on);
    srand(12345);
    for (int t = 0; t < 24; ++t) {
        for (int l = 0; l < 32; ++l) { w[l] = rand() % (t < 8 ? 32 : t < 16 ? 64 : 1400); on[l] = t % 4 == 3 ? (rand() % 8 != 0) : 1; }
        add("random", w, on);
    }
    for (size_t i = 0; i < W.size(); ++i) {
        cudaMemcpy(d_w, W[i].data(), 128, cudaMemcpyHostToDevice); cudaMemcpy(d_on, ON[i].data(), 128, cudaMemcpyHostToDevice);
        probe<<<1, 32>>>(d_w, d_on, d_out, 1000);
        cudaDeviceSynchronize();
        printf("launch %zu model %d per store : %s\n", i, model(W[i], ON[i]), name[i]);
    }
    return 0;
}
