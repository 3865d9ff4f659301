// cv::imwrite("*.png", 8-bit image) as the reference drivers use it for their results (AD-CensusV1/main.cpp:112-114,
// ASW/ASWeight.cpp:75-92, SAD/SADmain.cpp:96-98): a PNG writer for 8-bit grey and BGR images.
// Host code only.  The pixel data is exact; the container is not byte-identical to OpenCV's (which deflates with libpng /
// zlib): the image data goes into stored (uncompressed) deflate blocks, which every PNG reader accepts.
#include <cstdio>
#include <vector>

#include "common.cuh"

namespace {

uint32_t crc_table[256];
bool crc_ready = false;
void crc_init() {
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
    crc_ready = true;
}
uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0xffffffffu) {
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c;
}
void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}
void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
    put32(out, (uint32_t)data.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, crc32(out.data() + at, out.size() - at) ^ 0xffffffffu);
}

}  // namespace

extern "C" smt_status_t smt_write_png_u8(smt_ctx* ctx, const char* path, const uint8_t* img, int32_t rows, int32_t cols, int32_t channels,
                                         smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, path && img && rows > 0 && cols > 0 && (channels == 1 || channels == 3), SMT_ERR_INVALID_ARG,
                "smt_write_png_u8: bad argument (8-bit images with 1 or 3 channels)");
    const size_t rowb = (size_t)cols * channels, n = rowb * rows;
    std::vector<uint8_t> host;
    if (space == SMT_DEVICE) {
        DeviceGuard dev_guard(ctx->device);
        host.resize(n);
        SMT_CUDA(ctx, cudaMemcpyAsync(host.data(), img, n, cudaMemcpyDeviceToHost, ctx->stream));
        SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        img = host.data();
    }
    if (!crc_ready) crc_init();
    // scanlines: filter byte 0 + the row; OpenCV stores BGR, PNG wants RGB
    std::vector<uint8_t> raw;
    raw.reserve((rowb + 1) * rows);
    for (int i = 0; i < rows; ++i) {
        raw.push_back(0);
        const uint8_t* r = img + (size_t)i * rowb;
        if (channels == 1) {
            raw.insert(raw.end(), r, r + rowb);
        } else {
            for (int j = 0; j < cols; ++j) {
                raw.push_back(r[3 * j + 2]);
                raw.push_back(r[3 * j + 1]);
                raw.push_back(r[3 * j]);
            }
        }
    }
    // zlib stream of stored blocks
    std::vector<uint8_t> z;
    z.push_back(0x78);
    z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size();) {
        const size_t len = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + len == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(len & 0xff));
        z.push_back((uint8_t)(len >> 8));
        z.push_back((uint8_t)(~len & 0xff));
        z.push_back((uint8_t)((~len >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        for (size_t k = off; k < off + len; ++k) {  // Adler-32
            a = (a + raw[k]) % 65521u;
            b = (b + a) % 65521u;
        }
        off += len;
    }
    put32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)cols);
    put32(ihdr, (uint32_t)rows);
    ihdr.push_back(8);
    ihdr.push_back(channels == 1 ? 0 : 2);
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return smt_fail(ctx, SMT_ERR_INVALID_ARG, "smt_write_png_u8: cannot open %s", path);
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    const bool closed = std::fclose(f) == 0;
    if (!ok || !closed) return smt_fail(ctx, SMT_ERR_INVALID_ARG, "smt_write_png_u8: short write to %s", path);
    return SMT_OK;
}
