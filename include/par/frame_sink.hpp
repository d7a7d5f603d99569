// frame_sink.hpp — host-side writers for the frames the render path returns: the step AFTER the path.
// The reference hands its RGBA8 texture to SDL (/root/reference/src/alternative.cpp:774-788) and its README
// shows the result as an animated GIF (gif.gif); a headless host writes the same bytes to files instead:
//   write_ppm   one binary P6 file per frame (RGB, alpha dropped)
//   write_png   one PNG per frame (8-bit RGB, zlib "stored" blocks: no compression library needed, lossless)
//   GifWriter   one animated GIF89a for a frame sequence (LZW written here; a frame with <= 256 distinct colours
//               keeps them exactly, otherwise its colours are mapped to a 6x7x6 uniform palette)
// Header-only, no dependencies beyond the C++ standard library; frames are rows of par::Color-compatible
// 4-byte RGBA pixels with a caller-given row pitch in bytes (the locked-texture contract of
// alternative.cpp:774-783; pitch = width * 4 for packed frames).  Used by host/par_headless.cpp; tested on
// the CPU by tests/test_frame_sink.py.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace par {
namespace sink {

inline const unsigned char* row_of(const void* frame, size_t pitch, int y) {
    return static_cast<const unsigned char*>(frame) + static_cast<size_t>(y) * pitch;
}

// ---- PPM -------------------------------------------------------------------------------------------
inline bool write_ppm(const char* path, const void* rgba, int width, int height, size_t pitch) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    fprintf(fp, "P6\n%d %d\n255\n", width, height);
    std::vector<unsigned char> rgb(static_cast<size_t>(width) * 3);
    bool ok = true;
    for (int y = 0; y < height && ok; y++) {
        const unsigned char* src = row_of(rgba, pitch, y);
        for (int x = 0; x < width; x++) memcpy(&rgb[3 * static_cast<size_t>(x)], src + 4 * static_cast<size_t>(x), 3);
        ok = fwrite(rgb.data(), 1, rgb.size(), fp) == rgb.size();
    }
    return fclose(fp) == 0 && ok;
}

// ---- PNG -------------------------------------------------------------------------------------------
inline uint32_t crc32_update(uint32_t crc, const unsigned char* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}

inline void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    for (int s = 24; s >= 0; s -= 8) v.push_back(static_cast<unsigned char>(x >> s));
}

inline bool write_png_chunk(FILE* fp, const char type[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, static_cast<uint32_t>(data.size()));
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xffffffffu, reinterpret_cast<const unsigned char*>(type), 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<unsigned char> tail;
    put_be32(tail, crc);
    return fwrite(head.data(), 1, head.size(), fp) == head.size() &&
           (data.empty() || fwrite(data.data(), 1, data.size(), fp) == data.size()) &&
           fwrite(tail.data(), 1, tail.size(), fp) == tail.size();
}

inline bool write_png(const char* path, const void* rgba, int width, int height, size_t pitch) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = fwrite(sig, 1, 8, fp) == 8;
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, static_cast<uint32_t>(width));
    put_be32(ihdr, static_cast<uint32_t>(height));
    const unsigned char tail[5] = {8, 2, 0, 0, 0};  // 8 bits, colour type 2 (RGB), deflate, filter 0, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    ok = ok && write_png_chunk(fp, "IHDR", ihdr);
    // the raw image: a filter byte (0 = none) + RGB per row, wrapped into zlib stored blocks of <= 65535 bytes
    const size_t row = 1 + static_cast<size_t>(width) * 3;
    std::vector<unsigned char> raw(row * static_cast<size_t>(height));
    for (int y = 0; y < height; y++) {
        unsigned char* dst = &raw[row * static_cast<size_t>(y)];
        const unsigned char* src = row_of(rgba, pitch, y);
        dst[0] = 0;
        for (int x = 0; x < width; x++) memcpy(dst + 1 + 3 * static_cast<size_t>(x), src + 4 * static_cast<size_t>(x), 3);
    }
    uint32_t a = 1, b = 0;  // Adler-32 of the raw image
    for (size_t i = 0; i < raw.size();) {
        const size_t n = raw.size() - i < 5552 ? raw.size() - i : 5552;
        for (size_t k = 0; k < n; k++) {
            a += raw[i + k];
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
        i += n;
    }
    std::vector<unsigned char> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78);
    z.push_back(0x01);
    for (size_t i = 0; i < raw.size() || i == 0;) {
        const size_t n = raw.size() - i < 65535 ? raw.size() - i : 65535;
        z.push_back(i + n >= raw.size() ? 1 : 0);  // BFINAL, BTYPE = 00 (stored)
        z.push_back(static_cast<unsigned char>(n & 0xff));
        z.push_back(static_cast<unsigned char>(n >> 8));
        z.push_back(static_cast<unsigned char>(~n & 0xff));
        z.push_back(static_cast<unsigned char>((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + static_cast<long>(i), raw.begin() + static_cast<long>(i + n));
        i += n;
        if (n == 0) break;
    }
    put_be32(z, b << 16 | a);
    ok = ok && write_png_chunk(fp, "IDAT", z) && write_png_chunk(fp, "IEND", {});
    return fclose(fp) == 0 && ok;
}

// ---- animated GIF ----------------------------------------------------------------------------------
class GifWriter {
public:
    // delay_cs: display time of a frame in 1/100 s (the reference's gif.gif plays at about 30 frames/s: 3)
    GifWriter(const char* path, int width, int height, int delay_cs = 3)
        : fp_(fopen(path, "wb")), w_(width), h_(height), delay_(delay_cs) {
        if (!fp_) return;
        unsigned char head[13] = {'G', 'I', 'F', '8', '9', 'a'};
        head[6] = static_cast<unsigned char>(w_ & 0xff);
        head[7] = static_cast<unsigned char>(w_ >> 8);
        head[8] = static_cast<unsigned char>(h_ & 0xff);
        head[9] = static_cast<unsigned char>(h_ >> 8);
        head[10] = 0x70;  // no global colour table, 8 bits of colour resolution
        ok_ = fwrite(head, 1, 13, fp_) == 13;
        static const unsigned char loop[19] = {0x21, 0xff, 11, 'N', 'E', 'T', 'S', 'C', 'A', 'P', 'E', '2', '.', '0', 3, 1, 0, 0, 0};
        ok_ = ok_ && fwrite(loop, 1, 19, fp_) == 19;  // loop forever
    }
    GifWriter(const GifWriter&) = delete;
    GifWriter& operator=(const GifWriter&) = delete;
    ~GifWriter() { close(); }

    bool ok() const { return fp_ && ok_; }
    int frames() const { return frames_; }
    bool last_frame_exact() const { return exact_; }  // the last frame kept its colours (<= 256 of them)

    bool add_frame(const void* rgba, size_t pitch) {
        if (!ok()) return false;
        // palette: the frame's own colours when there are at most 256, else 6x7x6 uniform levels
        std::vector<unsigned char> index(static_cast<size_t>(w_) * h_);
        std::vector<uint32_t> colours;
        std::unordered_map<uint32_t, int> slot;
        exact_ = true;
        for (int y = 0; y < h_ && exact_; y++) {
            const unsigned char* src = row_of(rgba, pitch, y);
            for (int x = 0; x < w_; x++) {
                const uint32_t c = src[4 * x] | src[4 * x + 1] << 8 | src[4 * x + 2] << 16;
                auto it = slot.find(c);
                if (it == slot.end()) {
                    if (colours.size() == 256) {
                        exact_ = false;
                        break;
                    }
                    it = slot.emplace(c, static_cast<int>(colours.size())).first;
                    colours.push_back(c);
                }
                index[static_cast<size_t>(y) * w_ + x] = static_cast<unsigned char>(it->second);
            }
        }
        if (!exact_) {
            colours.clear();
            for (int r = 0; r < 6; r++)
                for (int g = 0; g < 7; g++)
                    for (int b = 0; b < 6; b++)
                        colours.push_back(static_cast<uint32_t>(r * 255 / 5) | static_cast<uint32_t>(g * 255 / 6) << 8 |
                                          static_cast<uint32_t>(b * 255 / 5) << 16);
            for (int y = 0; y < h_; y++) {
                const unsigned char* src = row_of(rgba, pitch, y);
                for (int x = 0; x < w_; x++) {
                    const int r = (src[4 * x] * 5 + 127) / 255, g = (src[4 * x + 1] * 6 + 127) / 255, b = (src[4 * x + 2] * 5 + 127) / 255;
                    index[static_cast<size_t>(y) * w_ + x] = static_cast<unsigned char>((r * 7 + g) * 6 + b);
                }
            }
        }
        int bits = 1;
        while ((1u << bits) < colours.size()) bits++;
        if (bits < 2) bits = 2;  // LZW minimum code size is 2 even for 2 colours
        const unsigned char gce[8] = {0x21, 0xf9, 4, 0x04, static_cast<unsigned char>(delay_ & 0xff),
                                      static_cast<unsigned char>(delay_ >> 8), 0, 0};
        ok_ = fwrite(gce, 1, 8, fp_) == 8;
        unsigned char desc[10] = {0x2c, 0, 0, 0, 0};
        desc[5] = static_cast<unsigned char>(w_ & 0xff);
        desc[6] = static_cast<unsigned char>(w_ >> 8);
        desc[7] = static_cast<unsigned char>(h_ & 0xff);
        desc[8] = static_cast<unsigned char>(h_ >> 8);
        desc[9] = static_cast<unsigned char>(0x80 | (bits - 1));  // local colour table of 2^bits entries
        ok_ = ok_ && fwrite(desc, 1, 10, fp_) == 10;
        std::vector<unsigned char> table(3u << bits, 0);
        for (size_t i = 0; i < colours.size(); i++) {
            table[3 * i] = static_cast<unsigned char>(colours[i] & 0xff);
            table[3 * i + 1] = static_cast<unsigned char>(colours[i] >> 8 & 0xff);
            table[3 * i + 2] = static_cast<unsigned char>(colours[i] >> 16 & 0xff);
        }
        ok_ = ok_ && fwrite(table.data(), 1, table.size(), fp_) == table.size();
        lzw(index, bits);
        frames_++;
        return ok_;
    }

    bool close() {
        if (!fp_) return false;
        const unsigned char trailer = 0x3b;
        ok_ = ok_ && fwrite(&trailer, 1, 1, fp_) == 1;
        ok_ = (fclose(fp_) == 0) && ok_;
        fp_ = nullptr;
        return ok_;
    }

private:
    // GIF-flavoured LZW: codes of variable width (min_bits + 1 .. 12), clear code when the table is full,
    // output packed LSB first into sub-blocks of at most 255 bytes.
    void lzw(const std::vector<unsigned char>& px, int min_bits) {
        const int clear = 1 << min_bits, eoi = clear + 1;
        const unsigned char mb = static_cast<unsigned char>(min_bits);
        ok_ = ok_ && fwrite(&mb, 1, 1, fp_) == 1;
        std::vector<unsigned char> out;
        uint32_t acc = 0;
        int n_acc = 0;
        auto emit = [&](int code, int width) {
            acc |= static_cast<uint32_t>(code) << n_acc;
            n_acc += width;
            while (n_acc >= 8) {
                out.push_back(static_cast<unsigned char>(acc & 0xff));
                acc >>= 8;
                n_acc -= 8;
            }
        };
        std::vector<int> next(4096 * 256, -1);  // (prefix code, byte) -> code; reset by index list
        std::vector<int> used;
        used.reserve(4096);
        int n_codes = eoi + 1, width = min_bits + 1;
        emit(clear, width);
        int cur = px.empty() ? -1 : px[0];
        for (size_t i = 1; i < px.size(); i++) {
            const int key = cur * 256 + px[i];
            if (next[key] >= 0) {
                cur = next[key];
                continue;
            }
            emit(cur, width);
            if (n_codes < 4096) {
                next[key] = n_codes++;
                used.push_back(key);
                if (n_codes > (1 << width) && width < 12) width++;
            } else {  // table full: start over
                emit(clear, width);
                for (int k : used) next[k] = -1;
                used.clear();
                n_codes = eoi + 1;
                width = min_bits + 1;
            }
            cur = px[i];
        }
        if (cur >= 0) {
            emit(cur, width);
            // the decoder adds one more table entry when it reads this code and may widen its codes by doing
            // so: the end code must be written at the width it will then read with
            if (n_codes < 4096 && ++n_codes > (1 << width) && width < 12) width++;
        }
        emit(eoi, width);
        if (n_acc > 0) out.push_back(static_cast<unsigned char>(acc & 0xff));
        for (size_t i = 0; i < out.size(); i += 255) {
            const unsigned char n = static_cast<unsigned char>(out.size() - i < 255 ? out.size() - i : 255);
            ok_ = ok_ && fwrite(&n, 1, 1, fp_) == 1 && fwrite(&out[i], 1, n, fp_) == n;
        }
        const unsigned char zero = 0;
        ok_ = ok_ && fwrite(&zero, 1, 1, fp_) == 1;
    }

    FILE* fp_;
    int w_, h_, delay_;
    int frames_ = 0;
    bool ok_ = false, exact_ = true;
};

}  // namespace sink
}  // namespace par
