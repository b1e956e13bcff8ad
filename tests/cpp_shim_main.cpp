// Exercises the header-only vs::Stabilizer shim exactly the way the reference's apps call the class
// (examples/vsg.cpp:1285, examples/file-capture.cpp:64): construct from Parameters, push frames,
// keep the raw frame while an empty Mat comes back, flush at the end.  Prints one CRC per output.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "video/RollCorrection.h"
#include "video/Stabilizer.h"

static uint32_t crc32(const uint8_t* p, size_t n, uint32_t c) {
    c = ~c;
    for (size_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    }
    return ~c;
}
static uint32_t mat_crc(const cv::Mat& m) {
    uint32_t c = 0;
    for (int y = 0; y < m.rows; ++y) c = crc32(m.data + (size_t)y * m.step, (size_t)m.cols * 3, c);
    return c;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s clip.raw width height frames [smoothing_radius]\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), n = std::atoi(argv[4]);
    vs::Stabilizer::Parameters params;
    params.smoothingRadius = argc > 5 ? std::atoi(argv[5]) : 5;
    const bool roll = argc > 6 && !std::strcmp(argv[6], "roll");     // the reference's pipeline order: roll correction, then stabilize (vsg.cpp:1272-1285)
    try {
        vs::Stabilizer stab(params);
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) { std::fprintf(stderr, "cannot open clip\n"); return 2; }
        for (int i = 0; i < n; ++i) {
            cv::Mat frame(h, w, CV_8UC3);
            if (std::fread(frame.data, 1, (size_t)w * h * 3, f) != (size_t)w * h * 3) return 2;
            if (roll) frame = vs::RollCorrection::autoCorrectRoll(frame);
            cv::Mat out = stab.stabilize(frame);
            if (!out.empty()) std::printf("%08x %d %d\n", mat_crc(out), out.cols, out.rows);
        }
        std::fclose(f);
        for (;;) {
            cv::Mat out = stab.flush();
            if (out.empty()) break;
            std::printf("%08x %d %d\n", mat_crc(out), out.cols, out.rows);
        }
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 3;
    }
    return 0;
}
