// Minimal PNG reader for the image source of VO::read_img: the reference loads KITTI's image_0 / image_1 frames with
// cv::imread(path, CV_LOAD_IMAGE_GRAYSCALE) (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:42-51);
// those files are 8-bit greyscale, non-interlaced PNGs.  Supported: colour type 0 (grey) at 8 bits; colour types 2 / 4
// / 6 at 8 bits are reduced to grey the way cv::imread does it (libpng's rgb_to_gray: (9797 R + 19234 G + 3737 B) >> 15).  16-bit,
// palette and interlaced files are rejected (return false).  Inflate comes from zlib.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace vslam {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// grey: w*h bytes, row-major
bool read_png_gray8(const std::string& path, std::vector<uint8_t>& grey, int& w, int& h) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    int depth = 0, ctype = -1, interlace = 0;
    w = h = 0;
    std::vector<uint8_t> idat;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return false;
        const uint8_t* data = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            w = (int)be32(data);
            h = (int)be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (w <= 0 || h <= 0 || depth != 8 || interlace != 0) return false;
    int ch;
    switch (ctype) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: return false;
    }
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) return false;
    std::vector<uint8_t> img(stride * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return false;
            }
            cur[x] = (uint8_t)v;
        }
    }
    grey.resize((size_t)w * h);
    if (ch == 1) {
        grey = img;
    } else if (ch == 2) {
        for (size_t i = 0; i < (size_t)w * h; ++i) grey[i] = img[2 * i];
    } else {  // libpng png_set_rgb_to_gray(0.299, 0.587) as cv::imread(GRAYSCALE) uses it: 15-bit truncated weights
        for (size_t i = 0; i < (size_t)w * h; ++i) {
            const uint8_t* p = &img[(size_t)ch * i];
            grey[i] = (uint8_t)((p[0] * 9797 + p[1] * 19234 + p[2] * 3737) >> 15);
        }
    }
    return true;
}

}  // namespace vslam

// test tap (ctypes): decode into a caller buffer of cap bytes; returns 0 on success
extern "C" int vslam_host_read_png(const char* path, uint8_t* out, int cap, int* w, int* h) {
    std::vector<uint8_t> g;
    int ww = 0, hh = 0;
    if (!path || !vslam::read_png_gray8(path, g, ww, hh)) return -1;
    if (w) *w = ww;
    if (h) *h = hh;
    if (!out || (size_t)cap < g.size()) return -2;
    std::memcpy(out, g.data(), g.size());
    return 0;
}
