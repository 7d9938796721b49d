// io_formats.cuh — streaming readers of the reference's wire formats,
// shared by the C ABI in io_formats.cu and by xrb_match_upload_ftr (match_api.cu).
//
// Format sources: src/utility/io_base.hpp:13-87 (raw little-endian dumps, NUL-terminated
// names), src/utility/io_feature.hpp:37-147 (ftr.bin, fp.bin), src/utility/io_ecim.cc:9-87,
// 145-232 (cameras.bin / images.bin / points3D.bin).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>

namespace xrb {

// Buffered binary file with bounds checks: every read and skip is validated against the file
// size, so a truncated or foreign file yields XRB_ERR_INVALID instead of garbage sizes.
class BinFile {
  public:
    ~BinFile() { close(); }
    bool open_read(const char *path);
    bool open_write(const char *path);
    bool close();  // false when the final flush failed (ENOSPC, EIO): writers fold it into their status
    bool read(void *dst, size_t bytes);
    bool skip(int64_t bytes);
    bool write(const void *src, size_t bytes);
    bool read_name(std::string *name, size_t max_len = 4096);  // up to and excluding the NUL
    int64_t tell() const { return pos_; }
    int64_t size() const { return size_; }
    const char *path() const { return path_.c_str(); }
    template <class T>
    bool get(T *v) {
        return read(v, sizeof(T));
    }
    template <class T>
    bool put(const T &v) {
        return write(&v, sizeof(T));
    }

  private:
    FILE *f_ = nullptr;
    int64_t pos_ = 0, size_ = 0;
    std::string path_;
};

// ftr.bin frame iterator: header() then either keypoints()+descriptors() or skip_body().
class FtrReader {
  public:
    int open(const char *path);  // XRB_OK or an error status (message set)
    int n_frames() const { return n_frames_; }
    // next frame: its name and point count; XRB_OK / error
    int header(std::string *name, int32_t *n_points);
    int keypoints(float *dst);          // n_points x 4 floats, or skipped when dst == nullptr
    int descriptors(uint8_t *dst);      // n_points x 128 bytes, or skipped when dst == nullptr

  private:
    BinFile f_;
    int32_t n_frames_ = 0, cur_points_ = 0;
};

}  // namespace xrb
