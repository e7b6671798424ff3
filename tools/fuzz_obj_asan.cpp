// ASan/UBSan harness: parseObjFast + parseObj over MUTATED (mostly invalid) OBJ texts -- the loader must answer with a
// mesh or an error, never crash and never touch a byte outside the text or its own buffers.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -ffp-contract=off -Idxrvoxelizer_b200/csrc tools/fuzz_obj_asan.cpp \
//       dxrvoxelizer_b200/csrc/obj_loader.cpp -o /tmp/fuzz_obj_asan -lpthread
//   python tools/fuzz_obj.py --seeds 0:200 && /tmp/fuzz_obj_asan /tmp/fuzz_obj/fuzz_*.obj
// (60 mutants per file: byte flips, deletions, insertions of digits / slashes / record letters, truncation, huge
// indices, short faces; each parsed with 1 and 3 threads from an exact-size heap copy without a terminator.)
#include "obj_loader.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
int main(int argc, char** argv)
{
    unsigned long long accepted = 0, rejected = 0, fast = 0;
    for (int a = 1; a < argc; ++a)
    {
        FILE* f = fopen(argv[a], "rb"); if (!f) continue;
        std::string base; char buf[65536]; size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) base.append(buf, n);
        fclose(f);
        std::mt19937_64 rng(a * 7919u);
        for (int it = 0; it < 60; ++it)
        {
            std::string t = base;
            const int muts = 1 + (int)(rng() % 6);
            for (int m = 0; m < muts && !t.empty(); ++m)
            {
                const size_t pos = rng() % t.size();
                switch (rng() % 7)
                {
                case 0: t[pos] = (char)(rng() & 0xff); break;
                case 1: t.erase(pos, 1 + rng() % 8); break;
                case 2: t.insert(pos, 1, " \t\n/-+.0123456789efvnt#"[rng() % 23]); break;
                case 3: t.insert(pos, "99999999999"); break;
                case 4: t.insert(pos, "\nf 1 2\n"); break;
                case 5: t.resize(pos); break;
                case 6: t.insert(pos, "/"); break;
                }
            }
            // exact-size heap copy (no terminator): reads past the end are caught
            char* raw = (char*)malloc(t.size() ? t.size() : 1);
            memcpy(raw, t.data(), t.size());
            for (unsigned threads : {1u, 3u})
            {
                dxrv::ObjMesh m; std::string err;
                if (dxrv::parseObjFast(raw, t.size(), m, err, threads)) { ++fast; ++accepted; }
                else if (err.empty())
                {
                    dxrv::ObjMesh m2; std::string e2;
                    if (dxrv::parseObj(raw, t.size(), m2, e2)) ++accepted; else ++rejected;
                }
                else ++rejected;
            }
            free(raw);
        }
    }
    printf("accepted=%llu (fast %llu) rejected=%llu\n", accepted, fast, rejected);
    return 0;
}
