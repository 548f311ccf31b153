// vh_image_io.h — OpenCV-free readers/writers for the frame layout SaveFrame uses
// (/root/reference/src/SaveFrame.cpp:120-218): depth/<id>.png (16-bit gray, millimetres), RGB/<id>.png|.ppm,
// tcw/<id>.txt (4x4 row-major text). PNG support is the subset those files need: non-interlaced, 8/16-bit,
// gray / RGB / RGBA, all five scanline filters; zlib does the inflate/deflate. JPEG needs OpenCV (not in this image).
#ifndef VH_IMAGE_IO_H_
#define VH_IMAGE_IO_H_

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace vhio {

struct Raster {
  int width = 0, height = 0, channels = 0, bits = 0;   // bits per sample: 8 or 16
  std::vector<uint8_t> data;                            // row-major; 16-bit samples in host byte order
};

inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool read_png(const std::string& path, Raster& img) {
  std::vector<uint8_t> file;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (!read_file(path, file) || file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) return false;
  std::vector<uint8_t> idat;
  int color_type = -1, interlace = 0;
  for (size_t pos = 8; pos + 12 <= file.size();) {
    const uint32_t len = be32(&file[pos]);
    const uint8_t* type = &file[pos + 4];
    if (pos + 12 + (size_t)len > file.size()) return false;
    const uint8_t* body = &file[pos + 8];
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      img.width = (int)be32(body); img.height = (int)be32(body + 4); img.bits = body[8]; color_type = body[9]; interlace = body[12];
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (interlace != 0 || (img.bits != 8 && img.bits != 16)) return false;
  img.channels = color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0;
  if (!img.channels || img.width <= 0 || img.height <= 0) return false;
  const size_t bpp = (size_t)img.channels * img.bits / 8, stride = bpp * img.width;
  std::vector<uint8_t> raw((stride + 1) * img.height);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return false;
  img.data.assign(stride * img.height, 0);
  for (int y = 0; y < img.height; y++) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* in = &raw[(stride + 1) * y + 1];
    uint8_t* out = &img.data[stride * y];
    const uint8_t* up = y ? out - stride : nullptr;
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? out[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                  pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: return false;
      }
      out[i] = (uint8_t)(in[i] + pred);
    }
  }
  if (img.bits == 16) {   // big-endian samples -> host order
    uint16_t* s = reinterpret_cast<uint16_t*>(img.data.data());
    for (size_t i = 0; i < img.data.size() / 2; i++) { const uint8_t* p = &img.data[2 * i]; s[i] = (uint16_t)((p[0] << 8) | p[1]); }
  }
  return true;
}

inline void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* body, size_t len) {
  const uint8_t l[4] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len};
  out.insert(out.end(), l, l + 4);
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (len) out.insert(out.end(), body, body + len);
  const uint32_t crc = (uint32_t)crc32(0, &out[start], (uInt)(4 + len));
  const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
  out.insert(out.end(), c, c + 4);
}

inline bool write_png(const std::string& path, const Raster& img) {
  const int ct = img.channels == 1 ? 0 : img.channels == 3 ? 2 : img.channels == 4 ? 6 : -1;
  if (ct < 0 || (img.bits != 8 && img.bits != 16)) return false;
  const size_t bpp = (size_t)img.channels * img.bits / 8, stride = bpp * img.width;
  std::vector<uint8_t> raw((stride + 1) * img.height);
  for (int y = 0; y < img.height; y++) {
    raw[(stride + 1) * y] = 0;
    uint8_t* dst = &raw[(stride + 1) * y + 1];
    const uint8_t* src = &img.data[stride * y];
    if (img.bits == 8) std::memcpy(dst, src, stride);
    else for (size_t i = 0; i < stride / 2; i++) { const uint16_t v = reinterpret_cast<const uint16_t*>(src)[i]; dst[2 * i] = (uint8_t)(v >> 8); dst[2 * i + 1] = (uint8_t)v; }
  }
  std::vector<uint8_t> z(compressBound((uLong)raw.size()));
  uLongf zl = (uLongf)z.size();
  if (compress2(z.data(), &zl, raw.data(), (uLong)raw.size(), 3) != Z_OK) return false;
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  uint8_t ihdr[13] = {(uint8_t)(img.width >> 24), (uint8_t)(img.width >> 16), (uint8_t)(img.width >> 8), (uint8_t)img.width,
                      (uint8_t)(img.height >> 24), (uint8_t)(img.height >> 16), (uint8_t)(img.height >> 8), (uint8_t)img.height,
                      (uint8_t)img.bits, (uint8_t)ct, 0, 0, 0};
  put_chunk(out, "IHDR", ihdr, 13);
  put_chunk(out, "IDAT", z.data(), zl);
  put_chunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

// binary PPM (P6) / PGM (P5), 8 or 16 bits
inline bool read_pnm(const std::string& path, Raster& img) {
  std::vector<uint8_t> file;
  if (!read_file(path, file) || file.size() < 7 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) return false;
  img.channels = file[1] == '6' ? 3 : 1;
  size_t pos = 2;
  int vals[3], got = 0;
  while (got < 3 && pos < file.size()) {
    while (pos < file.size() && (file[pos] == ' ' || file[pos] == '\n' || file[pos] == '\r' || file[pos] == '\t')) pos++;
    if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') pos++; continue; }
    int v = 0; bool any = false;
    while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') { v = v * 10 + (file[pos++] - '0'); any = true; }
    if (!any) return false;
    vals[got++] = v;
  }
  pos++;   // single whitespace after maxval
  img.width = vals[0]; img.height = vals[1]; img.bits = vals[2] > 255 ? 16 : 8;
  const size_t need = (size_t)img.width * img.height * img.channels * (img.bits / 8);
  if (got < 3 || pos + need > file.size()) return false;
  img.data.assign(file.begin() + pos, file.begin() + pos + need);
  if (img.bits == 16) {
    uint16_t* s = reinterpret_cast<uint16_t*>(img.data.data());
    for (size_t i = 0; i < need / 2; i++) { const uint8_t* p = &img.data[2 * i]; s[i] = (uint16_t)((p[0] << 8) | p[1]); }
  }
  return true;
}

// nearest-neighbour resize of an 8-bit raster (frameLoad resizes RGB to 640x480, SaveFrame.cpp:171)
inline Raster resize_nearest(const Raster& in, int w, int h) {
  if (in.width == w && in.height == h) return in;
  Raster out; out.width = w; out.height = h; out.channels = in.channels; out.bits = in.bits;
  const size_t px = (size_t)in.channels * in.bits / 8;
  out.data.resize(px * w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int sx = (int)(((long long)x * in.width) / w), sy = (int)(((long long)y * in.height) / h);
      std::memcpy(&out.data[px * ((size_t)y * w + x)], &in.data[px * ((size_t)sy * in.width + sx)], px);
    }
  return out;
}

}  // namespace vhio
#endif  // VH_IMAGE_IO_H_
