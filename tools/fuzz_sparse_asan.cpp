// ASan/UBSan harness for the host side of DXRV_FORMAT_SPARSE_BRICKS: what dxrv_sparse_decode does (csrc/api.cu), over
// valid blobs (checked against the dense grid they were made from) and over MUTATED blobs -- header fields, state words,
// payload bytes, truncation.  A blob comes from a file or a peer: the decoder must answer false or fill exactly the
// caller's buffer, never read past the blob or write past the grid.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -mavx2 -Idxrvoxelizer_b200/csrc tools/fuzz_sparse_asan.cpp \
//       dxrvoxelizer_b200/csrc/sparse_host.cpp dxrvoxelizer_b200/csrc/host_pool.cpp -o /tmp/fuzz_sparse_asan -lpthread
//   /tmp/fuzz_sparse_asan blob.bin bits.bin [blob.bin bits.bin ...]     (pairs written by the numpy encoder of tests/test_sparse.py)
#include "sparse_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

static std::vector<unsigned char> readFile(const char* p)
{
    std::vector<unsigned char> v;
    FILE* f = fopen(p, "rb");
    if (!f) return v;
    unsigned char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
    fclose(f);
    return v;
}

// the body of dxrv_sparse_decode; dst is an exact-size heap block
static int decode(const unsigned char* blobData, size_t blobBytes, std::vector<uint32_t>* out, unsigned delayUs)
{
    // exact-size, 64-byte aligned copy of the blob (reads past its end are caught)
    void* blob = aligned_alloc(64, (blobBytes + 63) / 64 * 64 ? (blobBytes + 63) / 64 * 64 : 64);
    memcpy(blob, blobData, blobBytes);
    dxrv::SparseBlobView v;
    int rc = -1;
    if (dxrv::sparseParse(blob, blobBytes, v))
    {
        const size_t words = (size_t)(v.z1 - v.z0) * v.N * v.P;
        if (words <= (size_t)1 << 28)
        {
            uint32_t* dst = (uint32_t*)aligned_alloc(64, (words * 4 + 63) / 64 * 64 ? (words * 4 + 63) / 64 * 64 : 64);
            memset(dst, 0xA5, words * 4);
            dxrv::hostFillBegin(dst, v.N, v.z0, v.z1 - v.z0);
            if (delayUs) std::this_thread::sleep_for(std::chrono::microseconds(delayUs));
            const bool published = dxrv::hostFillPublish(v);
            const bool filled = dxrv::hostFillWait();
            rc = published && filled ? 0 : -2;
            if (out) out->assign(dst, dst + words);
            free(dst);
        }
    }
    free(blob);
    return rc;
}

int main(int argc, char** argv)
{
    unsigned long long ok = 0, refused = 0, bad = 0;
    for (int a = 1; a + 1 < argc; a += 2)
    {
        const std::vector<unsigned char> blob = readFile(argv[a]), bits = readFile(argv[a + 1]);
        for (unsigned delay : {0u, 2000u})
        {
            std::vector<uint32_t> out;
            if (decode(blob.data(), blob.size(), &out, delay) != 0 || out.size() * 4 != bits.size() || memcmp(out.data(), bits.data(), bits.size()) != 0)
            { printf("VALID BLOB NOT DECODED: %s (delay %u)\n", argv[a], delay); ++bad; }
        }
        {   // the host encoder: exact-size buffers, the same bytes as the blob on file
            dxrv::SparseBlobView v;
            void* copy = aligned_alloc(64, (blob.size() + 63) / 64 * 64);
            memcpy(copy, blob.data(), blob.size());
            if (dxrv::sparseParse(copy, blob.size(), v))
            {
                uint32_t* dense = (uint32_t*)malloc(bits.size() ? bits.size() : 4);
                memcpy(dense, bits.data(), bits.size());
                unsigned char* out = (unsigned char*)malloc(blob.size());
                size_t n = 0;
                const bool ok = dxrv::sparseEncode(dense, v.N, v.z0, v.z1, out, blob.size(), n);
                if (!ok || n != blob.size() || memcmp(out, blob.data(), n) != 0) { printf("ENCODER DIFFERS: %s\n", argv[a]); ++bad; }
                size_t need = 0;
                if (dxrv::sparseEncode(dense, v.N, v.z0, v.z1, out, blob.size() - 1, need) || need != blob.size()) { printf("ENCODER CAPACITY CHECK: %s\n", argv[a]); ++bad; }
                free(out); free(dense);
            }
            free(copy);
        }
        std::mt19937_64 rng(a * 104729u);
        for (int it = 0; it < 400; ++it)
        {
            std::vector<unsigned char> t = blob;
            const int muts = 1 + (int)(rng() % 3);
            for (int m = 0; m < muts && t.size() >= 64; ++m)
            {
                switch (rng() % 6)
                {
                case 0: { const size_t w = 2 + rng() % 13; uint32_t x; memcpy(&x, &t[4 * w], 4);       // a header field: small step, bit flip or random
                          const unsigned k = rng() % 4; x = k == 0 ? x + 1 : k == 1 ? x - 1 : k == 2 ? x ^ (1u << (rng() % 32)) : (uint32_t)rng();
                          memcpy(&t[4 * w], &x, 4); break; }
                case 1: t[64 + rng() % (t.size() - 64 ? t.size() - 64 : 1)] ^= (unsigned char)(1u << (rng() % 8)); break;   // states / payload
                case 2: t.resize(rng() % t.size()); break;                                                                // truncated
                case 3: t.resize(t.size() + 1 + rng() % 256, 0xff); break;                                                // trailing bytes
                case 4: { const size_t p = 64 + rng() % 64; if (p < t.size()) t[p] = 0xff; break; }                      // state 3
                case 5: { const size_t p = rng() % t.size(); t[p] = (unsigned char)rng(); break; }
                }
            }
            const int rc = decode(t.data(), t.size(), nullptr, it % 4 == 0 ? 300u : 0u);
            if (rc == 0) ++ok; else ++refused;
        }
    }
    printf("mutants decoded=%llu refused=%llu, valid blobs failing=%llu\n", ok, refused, bad);
    return bad ? 1 : 0;
}
