// Known-answer vectors for curand_init(seed, subsequence, offset) (XORWOW), computed ON THE HOST by the CUDA toolkit's own
// curand_kernel.h (its host code path uses precalc_xorwow_*_host).  Output: JSON on stdout -> tests/golden/xorwow_init_kat.json
//   nvcc -o /tmp/make_xorwow_kat tools/make_xorwow_kat.cu && /tmp/make_xorwow_kat > tests/golden/xorwow_init_kat.json
#define QUALIFIERS static inline __host__ __device__
#include <cstdio>
#include <curand_kernel.h>

int main()
{
    const unsigned long long seeds[] = {0ull, 0xC0FFEEull, 12345ull, 0x123456789abcdefull, 31337ull};
    const unsigned long long subs[] = {0, 1, 2, 3, 7, 63, 64, 95, 239, 1199, 2399};
    const unsigned long long offs[] = {0, 1, 2, 5, 31, 127, 319, 1599, 3199};
    printf("{\"toolkit\": \"CUDA %d.%d curand_kernel.h, host path\", \"vectors\": [\n", CUDART_VERSION / 1000, (CUDART_VERSION % 1000) / 10);
    bool first = true;
    for (unsigned long long seed : seeds)
        for (unsigned long long sub : subs)
            for (unsigned long long off : offs) {
                curandStateXORWOW_t st;
                curand_init(seed, sub, off, &st);
                const unsigned v0 = st.v[0], v1 = st.v[1], v2 = st.v[2], v3 = st.v[3], v4 = st.v[4], d = st.d;
                const unsigned r0 = curand(&st), r1 = curand(&st);
                printf("%s[%llu, %llu, %llu, %u, %u, %u, %u, %u, %u, %u, %u]", first ? "" : ",\n", seed, sub, off, v0, v1, v2, v3, v4, d, r0, r1);
                first = false;
            }
    printf("\n]}\n");
    return 0;
}
