/* pt_io.cpp -- headless output: the reference's display transform + 8-bit PPM, and raw PFM.
 *
 * pt_write_ppm = SaveRender + SavePPM (host:3491-3518, 918-933) with the CPU twin of shader.frag:31-93
 * (host:949-1071): Bradford E->D65, XYZ->linear sRGB, max(0), tonemap {0 none, 1 Reinhard, 2 ACES fit,
 * 3 exp(-0.25/x)}, sRGB companding, byte = (char)(c * 255).  Buffer row 0 is the top scanline.
 * pt_write_pfm keeps the float data ("PF", little-endian, bottom-up by the format's convention -> rows flipped).
 * This is post-processing, outside the parity-gated kernel: it uses libm.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "pt_abi.h"

namespace {

void xyz_to_linear_srgb(const float* xyz, float* rgb) {
    /* v * M with GLSL/glm column-major constructors: component j = dot(v, column j) */
    static const float E2D65[9] = {0.9531874f, -0.0265906f, 0.0238731f, -0.0382467f, 1.0288406f, 0.0094060f,
                                   0.0026068f, -0.0030332f, 1.0892565f};
    static const float X2R[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f,
                                 0.0556434f, -0.2040259f, 1.0572252f};
    float d[3];
    for (int j = 0; j < 3; j++) d[j] = xyz[0] * E2D65[3 * j] + xyz[1] * E2D65[3 * j + 1] + xyz[2] * E2D65[3 * j + 2];
    for (int j = 0; j < 3; j++) rgb[j] = d[0] * X2R[3 * j] + d[1] * X2R[3 * j + 1] + d[2] * X2R[3 * j + 2];
}

float tonemap_one(float x, int tonemap) {
    if (tonemap == 1) return x / (1.0f + x);
    if (tonemap == 2) return x * (2.51f * x + 0.03f) / (x * (2.43f * x + 0.59f) + 0.14f);
    if (tonemap == 3) return expf(-0.25f / x);
    return x;
}

float srgb_companding(float x) {
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (x <= 0.0031308f) ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}

}  // namespace

extern "C" {

int pt_write_ppm(const char* path, const float* rgba, int width, int height, int tonemap) {
    if (!path || !rgba || width <= 0 || height <= 0) return PT_ERR_ARG;
    std::vector<unsigned char> bytes((size_t)width * (size_t)height * 3);
    for (size_t i = 0; i < (size_t)width * (size_t)height; i++) {
        float rgb[3];
        xyz_to_linear_srgb(rgba + 4 * i, rgb);
        for (int k = 0; k < 3; k++) {
            float v = rgb[k] < 0.0f ? 0.0f : rgb[k]; /* glm::max(x, 0): NaN stays NaN -> companding clamps */
            v = srgb_companding(tonemap_one(v, tonemap));
            if (!(v == v)) v = 0.0f;
            bytes[3 * i + k] = (unsigned char)(int)(v * 255.0);
        }
    }
    FILE* fp = fopen(path, "wb");
    if (!fp) return PT_ERR_IO;
    fprintf(fp, "P6\n%d\n%d\n255\n", width, height);
    const size_t w = fwrite(bytes.data(), 1, bytes.size(), fp);
    const int rc = (w == bytes.size() && fclose(fp) == 0) ? PT_OK : PT_ERR_IO;
    return rc;
}

int pt_write_pfm(const char* path, const float* rgba, int width, int height, int to_rgb) {
    if (!path || !rgba || width <= 0 || height <= 0) return PT_ERR_ARG;
    FILE* fp = fopen(path, "wb");
    if (!fp) return PT_ERR_IO;
    fprintf(fp, "PF\n%d %d\n-1.0\n", width, height);
    std::vector<float> row((size_t)width * 3);
    bool ok = true;
    for (int y = height - 1; y >= 0 && ok; y--) { /* PFM stores the bottom scanline first */
        for (int x = 0; x < width; x++) {
            const float* t = rgba + 4 * ((size_t)x + (size_t)width * (size_t)y);
            if (to_rgb) xyz_to_linear_srgb(t, &row[3 * (size_t)x]);
            else { row[3 * (size_t)x] = t[0]; row[3 * (size_t)x + 1] = t[1]; row[3 * (size_t)x + 2] = t[2]; }
        }
        ok = fwrite(row.data(), sizeof(float), row.size(), fp) == row.size();
    }
    if (fclose(fp) != 0) ok = false;
    return ok ? PT_OK : PT_ERR_IO;
}

/* OpenEXR 2 single-part scanline file, NO_COMPRESSION, three FLOAT channels: R,G,B (linear sRGB through the
 * reference's Bradford + XYZ->sRGB matrices) when to_rgb != 0, else X,Y,Z (the raw accumulation buffer).
 * Row 0 of the buffer is the top scanline = EXR's y = 0 (INCREASING_Y).  Written without any library: header
 * attributes, one offset per scanline, then per scanline {y, byte count, channel rows in alphabetical order}. */
int pt_write_exr(const char* path, const float* rgba, int width, int height, int to_rgb) {
    if (!path || !rgba || width <= 0 || height <= 0) return PT_ERR_ARG;
    std::string h;
    auto put32 = [&](std::string& o, uint32_t v) { for (int i = 0; i < 4; i++) o.push_back((char)((v >> (8 * i)) & 0xff)); };
    auto putf = [&](std::string& o, float f) { uint32_t u; memcpy(&u, &f, 4); put32(o, u); };
    auto attr = [&](const char* name, const char* type, const std::string& value) {
        h += name; h.push_back('\0'); h += type; h.push_back('\0');
        put32(h, (uint32_t)value.size());
        h += value;
    };
    put32(h, 20000630u); /* magic */
    put32(h, 2u);        /* version 2, single-part scanline, no long names */
    const char* names = to_rgb ? "BGR" : "XYZ"; /* alphabetical */
    std::string ch;
    for (int c = 0; c < 3; c++) {
        ch.push_back(names[c]); ch.push_back('\0');
        put32(ch, 2u);                      /* FLOAT */
        ch.push_back(0); ch.push_back(0); ch.push_back(0); ch.push_back(0); /* pLinear + reserved */
        put32(ch, 1u); put32(ch, 1u);       /* x/y sampling */
    }
    ch.push_back('\0');
    attr("channels", "chlist", ch);
    attr("compression", "compression", std::string(1, '\0'));
    std::string win;
    put32(win, 0u); put32(win, 0u); put32(win, (uint32_t)(width - 1)); put32(win, (uint32_t)(height - 1));
    attr("dataWindow", "box2i", win);
    attr("displayWindow", "box2i", win);
    attr("lineOrder", "lineOrder", std::string(1, '\0'));
    std::string one; putf(one, 1.0f);
    attr("pixelAspectRatio", "float", one);
    std::string c2; putf(c2, 0.0f); putf(c2, 0.0f);
    attr("screenWindowCenter", "v2f", c2);
    attr("screenWindowWidth", "float", one);
    h.push_back('\0');
    const uint64_t row_bytes = (uint64_t)width * 3 * 4, block = 8 + row_bytes;
    const uint64_t first = h.size() + (uint64_t)height * 8;
    FILE* fp = fopen(path, "wb");
    if (!fp) return PT_ERR_IO;
    bool ok = fwrite(h.data(), 1, h.size(), fp) == h.size();
    for (int y = 0; y < height && ok; y++) {
        const uint64_t off = first + (uint64_t)y * block;
        unsigned char b[8];
        for (int i = 0; i < 8; i++) b[i] = (unsigned char)((off >> (8 * i)) & 0xff);
        ok = fwrite(b, 1, 8, fp) == 8;
    }
    std::vector<float> row((size_t)width * 3);
    for (int y = 0; y < height && ok; y++) {
        std::string hd;
        put32(hd, (uint32_t)y); put32(hd, (uint32_t)row_bytes);
        ok = fwrite(hd.data(), 1, 8, fp) == 8;
        for (int x = 0; x < width; x++) {
            const float* t = rgba + 4 * ((size_t)x + (size_t)width * (size_t)y);
            float v[3] = {t[0], t[1], t[2]};
            if (to_rgb) { xyz_to_linear_srgb(t, v); const float r = v[0]; v[0] = v[2]; v[2] = r; } /* B, G, R */
            row[(size_t)x] = v[0]; row[(size_t)width + x] = v[1]; row[2 * (size_t)width + x] = v[2];
        }
        ok = ok && fwrite(row.data(), 4, row.size(), fp) == row.size();
    }
    if (fclose(fp) != 0) ok = false;
    return ok ? PT_OK : PT_ERR_IO;
}

/* Reads back what pt_write_pfm(..., to_rgb = 0) wrote: XYZ into the rgb of an RGBA buffer, w = 1 (checkpoints) */
int pt_read_pfm(const char* path, float* rgba, int width, int height) {
    if (!path || !rgba || width <= 0 || height <= 0) return PT_ERR_ARG;
    FILE* fp = fopen(path, "rb");
    if (!fp) return PT_ERR_IO;
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0;
    float scale = 0.0f;
    if (fscanf(fp, "%2s %d %d %f", magic, &w, &h, &scale) != 4 || magic[0] != 'P' || magic[1] != 'F' || w != width || h != height ||
        !(scale < 0.0f)) { /* little-endian files have a negative scale */
        fclose(fp);
        return PT_ERR_IO;
    }
    fgetc(fp); /* the single whitespace after the header */
    std::vector<float> row((size_t)width * 3);
    bool ok = true;
    for (int y = height - 1; y >= 0 && ok; y--) {
        ok = fread(row.data(), sizeof(float), row.size(), fp) == row.size();
        for (int x = 0; x < width && ok; x++) {
            float* t = rgba + 4 * ((size_t)x + (size_t)width * (size_t)y);
            t[0] = row[3 * (size_t)x]; t[1] = row[3 * (size_t)x + 1]; t[2] = row[3 * (size_t)x + 2]; t[3] = 1.0f;
        }
    }
    fclose(fp);
    return ok ? PT_OK : PT_ERR_IO;
}

} /* extern "C" */
