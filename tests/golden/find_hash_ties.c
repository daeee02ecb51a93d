/* Generator of tests/golden/hash_ties.json (test tooling, not product code).
 *
 * FreqFilter.add (S/data/FreqFilter.scala:31-32) stores `if (x.hashCode < rcx.hashCode) x else rcx`: when the two 32-bit
 * hashes TIE and x != rc(x), a read of x stores rc(x) and a read of rc(x) stores x -- both orientations of one k-mer end
 * up in the table.  Such k-mers are rare (2^-32 per k-mer) but they take paths of their own: the insert's
 * tie rule, contains() probing both orientations, the primary / secondary orientation of Graph.buildGraph.  This
 * program finds some by brute force so that the parity tests can exercise those paths with real reads.
 *
 * hash = Long.## of scala-library 2.9.1 on the 2-bit packed long (base i at bits 2i, A0 G1 C2 T3; variant 291) or of
 * scala >= 2.10 (variant 210).  Where ties exist: k <= 16 has none under either formula (the hash is a bijection of the
 * value).  ODD k in 17..31 has none under 291 either: the high word has 2k - 32 bits, so the bases k-16 .. 15 of the hash are the
 * bases of the k-mer itself, the middle base (k-1)/2 is among them, and rc(x) holds the COMPLEMENT of x's middle base there.
 * So only even k >= 18 can tie (x[i] ^ x[16+i] = x[k-17-i] ^ x[k-1-i] on the folded bases, x[i] = comp(x[k-1-i]) on the others:
 * about 2^-16 of all k-mers).
 *
 *   gcc -O2 -fopenmp -o /tmp/find_hash_ties tests/golden/find_hash_ties.c && /tmp/find_hash_ties > tests/golden/hash_ties.json
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>

static uint64_t rc(uint64_t x, int k)
{
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((x & 3) ^ 3);
        x >>= 2;
    }
    return r;
}
static int32_t h291(uint64_t v) { return (int32_t)((uint32_t)v ^ (uint32_t)(v >> 32)); }
static int32_t h210(uint64_t v)
{
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return (int32_t)(lo ^ (hi + (lo >> 31)));
}
static uint64_t splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

#define WANT 3
static void search(int k, int variant, int first)
{
    uint64_t found[WANT];
    int n = 0;
    const uint64_t mask = (1ull << (2 * k)) - 1;
#pragma omp parallel
    {
        uint64_t s = 0x1234567ull * (uint64_t)(k * 1000 + variant) + 0x9999ull * (uint64_t)omp_get_thread_num();
        for (;;) {
            int done;
#pragma omp atomic read
            done = n;
            if (done >= WANT) break;
            for (int it = 0; it < (1 << 20); it++) {
                const uint64_t x = splitmix(&s) & mask, r = rc(x, k);
                const int tie = variant == 291 ? h291(x) == h291(r) : h210(x) == h210(r);
                if (tie && x != r) {
#pragma omp critical
                    if (n < WANT) found[n++] = x;
                }
            }
        }
    }
    printf("%s\n  {\"k\": %d, \"variant\": %d, \"kmers\": [", first ? "" : ",", k, variant);
    for (int i = 0; i < WANT; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)found[i]);
    printf("]}");
    fflush(stdout);
}

int main(void)
{
    const int ks[] = { 18, 22, 26, 30 };
    printf("[");
    int first = 1;
    for (int i = 0; i < 4; i++) { search(ks[i], 291, first); first = 0; }
    search(30, 210, 0);
    printf("\n]\n");
    return 0;
}
