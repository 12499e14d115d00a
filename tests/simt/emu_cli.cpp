// emu_cli.cpp — tiny driver for the SIMT-emulator build (TEST INFRASTRUCTURE ONLY): encodes a file through the
// C ABI of liborz_b200_emu.so and writes the stream, so that emulator runs (timing model, sanitizers, race shaking)
// do not need Python.  usage: emu_cli LEVEL in out
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/orz_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s LEVEL in out\n", argv[0]);
        return 2;
    }
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 2;
    std::vector<uint8_t> in;
    uint8_t buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    orz_lzcfg cfg;
    if (orz_level_cfg(atoi(argv[1]), &cfg)) return 2;
    std::vector<uint8_t> out(orz_encode_bound(in.size()));
    size_t n = 0;
    const int rc = orz_encode_stream(0, &cfg, in.data(), in.size(), out.data(), out.size(), &n);
    if (rc) {
        fprintf(stderr, "orz_encode_stream: %s\n", orz_strerror(rc));
        return 1;
    }
    f = fopen(argv[3], "wb");
    fwrite(out.data(), 1, n, f);
    fclose(f);
    fprintf(stderr, "%zu -> %zu bytes\n", in.size(), n);
    return 0;
}
