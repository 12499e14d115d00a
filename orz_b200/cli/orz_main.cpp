// orz_main.cpp — command line front end with the interface of richox/orz's src/main.rs, over the C ABI of
// include/orz_b200.h.  Flags, defaults, exit behaviour and error strings follow main.rs:21-52,97-113; the
// statistics printed unless --silent follow progress.rs:62-98 (whole-stream figures: the codec runs on the GPU,
// so there is no per-block progress line).  No codec logic lives here.
#include "orz_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool read_all(FILE *f, std::vector<uint8_t> &buf) {
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    return !ferror(f);
}

static int usage() {
    fprintf(stderr, "orz: an optimized ROLZ data compressor (B200 build)\n\n"
                    "Usage: orz encode [-s|--silent] [-l|--level <0..2>] [--segment-size BYTES] [IPATH] [OPATH]\n"
                    "       orz decode [-s|--silent] [IPATH] [OPATH]\n");
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 2) return usage();
    const std::string cmd = argv[1];
    if (cmd != "encode" && cmd != "decode") return usage();
    const bool enc = cmd == "encode";
    bool silent = false;
    int level = 2; // main.rs:30 default_value = "2"
    size_t segment_size = 0; // extension: > 0 writes the segment container of include/orz_b200.h instead of one orz stream
    std::vector<const char *> paths;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-s" || a == "--silent") silent = true;
        else if (enc && (a == "-l" || a == "--level")) {
            if (++i >= argc) return usage();
            level = atoi(argv[i]);
        } else if (enc && a == "--segment-size") {
            if (++i >= argc) return usage();
            segment_size = (size_t)strtoull(argv[i], nullptr, 10);
        } else if (enc && a.rfind("-l", 0) == 0 && a.size() > 2) level = atoi(a.c_str() + 2);
        else if (enc && a.rfind("--level=", 0) == 0) level = atoi(a.c_str() + 8);
        else if (a == "-h" || a == "--help") return usage(), 0;
        else if (a.size() > 1 && a[0] == '-') return usage();
        else paths.push_back(argv[i]);
    }
    if (paths.size() > 2) return usage();
    orz_lzcfg cfg;
    if (enc && orz_level_cfg(level, &cfg) != ORZ_OK) { // main.rs:101
        fprintf(stderr, "Error: \"invalid level: %d\"\n", level);
        return 1;
    }
    FILE *fi = paths.size() >= 1 ? fopen(paths[0], "rb") : stdin; // main.rs:72-77
    if (!fi) {
        fprintf(stderr, "Error: cannot open %s\n", paths[0]);
        return 1;
    }
    std::vector<uint8_t> in;
    if (!read_all(fi, in)) {
        fprintf(stderr, "Error: \"%s failed: read error\"\n", enc ? "encoding" : "decoding");
        return 1;
    }
    if (fi != stdin) fclose(fi);

    const auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> out;
    size_t out_len = 0;
    int rc;
    size_t raw_len = 0;
    if (enc && segment_size > 0) {
        out.resize(orz_segmented_bound(in.size(), segment_size));
        rc = orz_encode_segmented(0, &cfg, level, in.data(), in.size(), segment_size, out.data(), out.size(), &out_len);
    } else if (enc) {
        out.resize(orz_encode_bound(in.size()));
        rc = orz_encode_stream(0, &cfg, in.data(), in.size(), out.data(), out.size(), &out_len);
    } else if (orz_segmented_raw_len(in.data(), in.size(), &raw_len) == ORZ_OK) { // "ORZS" container
        out.resize(raw_len + 1);
        rc = orz_decode_segmented(0, in.data(), in.size(), out.data(), raw_len, &out_len);
    } else {
        size_t cap = in.size() * 8 + (1 << 20);
        for (;;) { // the orz stream carries no size field: grow until it fits
            out.resize(cap);
            rc = orz_decode_stream(0, in.data(), in.size(), out.data(), out.size(), &out_len);
            if (rc != ORZ_ENOSPC || cap > ((size_t)1 << 40)) break;
            cap *= 4;
        }
    }
    if (rc != ORZ_OK) { // main.rs:105,113
        fprintf(stderr, "Error: \"%s failed: %s\"\n", enc ? "encoding" : "decoding", orz_strerror(rc));
        return 1;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FILE *fo = paths.size() >= 2 ? fopen(paths[1], "wb") : stdout; // main.rs:78-83
    if (!fo || fwrite(out.data(), 1, out_len, fo) != out_len) {
        fprintf(stderr, "Error: \"%s failed: write error\"\n", enc ? "encoding" : "decoding");
        return 1;
    }
    if (fo != stdout) fclose(fo);

    if (!silent) { // progress.rs:62-98
        const size_t ibs = in.size(), obs = out_len;
        fprintf(stderr, "statistics:\n");
        if (enc) fprintf(stderr, "  size:  %zu bytes => %zu bytes\n", ibs, obs);
        else fprintf(stderr, "  size:  %zu bytes <= %zu bytes\n", obs, ibs);
        const double ratio = enc ? obs * 100.0 / (ibs ? ibs : 1) : ibs * 100.0 / (obs ? obs : 1);
        fprintf(stderr, "  ratio: %.2f%%\n", ratio);
        fprintf(stderr, "  speed: %.3f MB/s\n", (enc ? ibs : obs) / (sec * 1e6));
        fprintf(stderr, "  time:  %.3f sec\n", sec);
    }
    return 0;
}
