// io_formats.cu — the reference's wire formats as flat arrays (host code).
//
// ftr.bin / fp.bin   src/utility/io_feature.hpp:37-147
// cameras.bin / images.bin / points3D.bin   src/utility/io_ecim.cc:9-87 (readers), 145-232 (writers)
// primitives         src/utility/io_base.hpp:13-87
// Every reader is two-pass (scan sizes, then fill caller-owned flat arrays): the layouts the two
// hot paths consume (packed descriptor block + row offsets, xrb_ba_problem) come straight off the
// file, without the reference's per-frame / per-track objects in between.
#include <sys/stat.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "io_formats.cuh"

namespace xrb {

// ---- BinFile -------------------------------------------------------------------------------
bool BinFile::open_read(const char *path) {
    close();
    path_ = path ? path : "";
    f_ = path ? fopen(path, "rb") : nullptr;
    if (!f_) return false;
    setvbuf(f_, nullptr, _IOFBF, 1 << 20);
    if (fseeko(f_, 0, SEEK_END) != 0) return false;
    size_ = (int64_t)ftello(f_);
    if (fseeko(f_, 0, SEEK_SET) != 0) return false;
    pos_ = 0;
    return true;
}
bool BinFile::open_write(const char *path) {
    close();
    path_ = path ? path : "";
    f_ = path ? fopen(path, "wb") : nullptr;
    if (!f_) return false;
    setvbuf(f_, nullptr, _IOFBF, 1 << 20);
    pos_ = size_ = 0;
    return true;
}
bool BinFile::close() {
    const bool ok = !f_ || fclose(f_) == 0;
    f_ = nullptr;
    return ok;
}
bool BinFile::read(void *dst, size_t bytes) {
    if (!f_ || pos_ + (int64_t)bytes > size_) return false;
    if (bytes && fread(dst, 1, bytes, f_) != bytes) return false;
    pos_ += (int64_t)bytes;
    return true;
}
bool BinFile::skip(int64_t bytes) {
    if (!f_ || bytes < 0 || pos_ + bytes > size_) return false;
    if (bytes && fseeko(f_, (off_t)bytes, SEEK_CUR) != 0) return false;
    pos_ += bytes;
    return true;
}
bool BinFile::write(const void *src, size_t bytes) {
    if (!f_) return false;
    if (bytes && fwrite(src, 1, bytes, f_) != bytes) return false;
    pos_ += (int64_t)bytes;
    return true;
}
bool BinFile::read_name(std::string *name, size_t max_len) {  // io_base.hpp:73-82
    if (name) name->clear();
    for (size_t n = 0; n <= max_len; ++n) {
        if (!f_ || pos_ >= size_) return false;
        const int c = fgetc(f_);
        if (c == EOF) return false;
        ++pos_;
        if (c == 0) return true;
        if (name) name->push_back((char)c);
    }
    return false;  // no terminator within max_len: not a name
}

namespace {

int bad_file(const BinFile &f, const char *what) {
    set_error("%s: %s (offset %lld of %lld bytes)", f.path(), what, (long long)f.tell(), (long long)f.size());
    return XRB_ERR_INVALID;
}
int cannot_open(const char *path) {
    set_error("cannot open %s", path ? path : "(null)");
    return XRB_ERR_INVALID;
}

constexpr int kCamParams[5] = {3, 4, 4, 5, 8};  // camera_model.hpp:95,114,133,157,181

}  // namespace

// ---- FtrReader -----------------------------------------------------------------------------
int FtrReader::open(const char *path) {
    if (!f_.open_read(path)) return cannot_open(path);
    if (!f_.get(&n_frames_) || n_frames_ < 0) return bad_file(f_, "ftr.bin: bad frame count");
    return XRB_OK;
}
int FtrReader::header(std::string *name, int32_t *n_points) {
    if (!f_.read_name(name)) return bad_file(f_, "ftr.bin: unterminated frame name");
    if (!f_.get(&cur_points_) || cur_points_ < 0 || cur_points_ > 1000000)  // io_feature.hpp:61
        return bad_file(f_, "ftr.bin: point count outside [0, 1e6]");
    *n_points = cur_points_;
    return XRB_OK;
}
int FtrReader::keypoints(float *dst) {
    const int64_t bytes = (int64_t)cur_points_ * 16;
    if (!(dst ? f_.read(dst, (size_t)bytes) : f_.skip(bytes))) return bad_file(f_, "ftr.bin: truncated keypoints");
    return XRB_OK;
}
int FtrReader::descriptors(uint8_t *dst) {
    const int64_t bytes = (int64_t)cur_points_ * 128;
    if (!(dst ? f_.read(dst, (size_t)bytes) : f_.skip(bytes))) return bad_file(f_, "ftr.bin: truncated descriptors");
    return XRB_OK;
}

}  // namespace xrb

using namespace xrb;

namespace {
struct FpPairHeader {
    int32_t id1, id2;
    uint64_t n_matches;
};
// walks the file; for each pair calls fn(header, file positioned at its matches), which must
// consume exactly the pair's body
template <class Fn>
int fp_walk(const char *path, Fn &&fn) {
    BinFile f;
    if (!f.open_read(path)) return cannot_open(path);
    uint64_t n = 0;
    if (!f.get(&n) || n > (uint64_t)f.size()) return bad_file(f, "fp.bin: bad pair count");
    for (uint64_t p = 0; p < n; ++p) {
        FpPairHeader h;
        if (!f.get(&h.id1) || !f.get(&h.id2) || !f.get(&h.n_matches) || h.n_matches > (uint64_t)f.size() / 16)
            return bad_file(f, "fp.bin: bad pair header");
        const int rc = fn(h, f);
        if (rc) return rc;
    }
    return XRB_OK;
}
}  // namespace

namespace {

std::string join(const char *dir, const char *file) { return std::string(dir ? dir : "") + file; }

// mkdir -p of the directory part of `dir` (the path prefix callers pass ends with '/')
void make_dirs(const char *dir) {
    std::string d(dir ? dir : "");
    for (size_t i = 1; i <= d.size(); ++i)
        if (i == d.size() || d[i] == '/') {
            const std::string sub = d.substr(0, i);
            if (!sub.empty() && sub != "/" && sub.back() != '/') mkdir(sub.c_str(), 0777);
        }
}

// points3D.bin: fn(index, id, xyz) per track, in file order
template <class Fn>
int walk_points(const std::string &path, Fn &&fn, int64_t *n_out) {
    BinFile f;
    if (!f.open_read(path.c_str())) return cannot_open(path.c_str());
    uint64_t n = 0;
    if (!f.get(&n) || n > (uint64_t)f.size() / 43) return bad_file(f, "points3D.bin: bad track count");
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t id = 0, n_obs = 0;
        double xyz[3], err;
        unsigned char rgb[3];
        if (!f.get(&id) || !f.read(xyz, 24) || !f.read(rgb, 3) || !f.get(&err) || !f.get(&n_obs) ||
            n_obs > (uint64_t)f.size() / 8 || !f.skip((int64_t)n_obs * 8))
            return bad_file(f, "points3D.bin: truncated track");
        fn((int64_t)i, id, xyz);
    }
    *n_out = (int64_t)n;
    return XRB_OK;
}

struct FrameHeader {
    uint32_t id, camera_id;
    double q[4], t[3];  // q as stored: w, x, y, z (io_ecim.cc:38-41)
    uint64_t n_p2d;
};
// images.bin: frame(index, header) then p2d(index, k, x, y, track_id) for each 2-D point
template <class FrameFn, class P2dFn>
int walk_images(const std::string &path, FrameFn &&on_frame, P2dFn &&on_p2d, int64_t *n_out) {
    BinFile f;
    if (!f.open_read(path.c_str())) return cannot_open(path.c_str());
    uint64_t n = 0;
    if (!f.get(&n) || n > (uint64_t)f.size() / 73) return bad_file(f, "images.bin: bad frame count");
    std::vector<unsigned char> buf;
    for (uint64_t i = 0; i < n; ++i) {
        FrameHeader h;
        if (!f.get(&h.id) || !f.read(h.q, 32) || !f.read(h.t, 24) || !f.get(&h.camera_id) || !f.read_name(nullptr) ||
            !f.get(&h.n_p2d) || h.n_p2d > (uint64_t)f.size() / 24)
            return bad_file(f, "images.bin: truncated frame header");
        const int rc = on_frame((int64_t)i, h);
        if (rc) return rc;
        buf.resize((size_t)h.n_p2d * 24);
        if (!f.read(buf.data(), buf.size())) return bad_file(f, "images.bin: truncated 2-D points");
        for (uint64_t k = 0; k < h.n_p2d; ++k) {
            double xy[2];
            uint64_t tid;
            memcpy(xy, buf.data() + 24 * k, 16);
            memcpy(&tid, buf.data() + 24 * k + 16, 8);
            on_p2d((int64_t)i, (int64_t)k, xy[0], xy[1], tid);
        }
    }
    *n_out = (int64_t)n;
    return XRB_OK;
}

}  // namespace

extern "C" {

// ---- ftr.bin -------------------------------------------------------------------------------
int xrb_ftr_scan(const char *path, int32_t *n_frames, int64_t *total_points, int64_t *names_bytes) {
    if (!path || !n_frames || !total_points) {
        set_error("ftr_scan: bad arguments");
        return XRB_ERR_INVALID;
    }
    FtrReader r;
    int rc = r.open(path);
    if (rc) return rc;
    int64_t total = 0, nb = 0;
    std::string name;
    for (int i = 0; i < r.n_frames(); ++i) {
        int32_t np = 0;
        if ((rc = r.header(&name, &np))) return rc;
        if ((rc = r.keypoints(nullptr))) return rc;
        if ((rc = r.descriptors(nullptr))) return rc;
        total += np, nb += (int64_t)name.size() + 1;
    }
    *n_frames = r.n_frames(), *total_points = total;
    if (names_bytes) *names_bytes = nb;
    return XRB_OK;
}

int xrb_ftr_read(const char *path, int32_t n_frames, int64_t *row_offsets, uint8_t *desc_block,
                 float *keypoints, char *names, int64_t *name_offsets) {
    if (!path || n_frames < 0 || !row_offsets) {
        set_error("ftr_read: bad arguments");
        return XRB_ERR_INVALID;
    }
    FtrReader r;
    int rc = r.open(path);
    if (rc) return rc;
    if (r.n_frames() != n_frames) {
        set_error("%s holds %d frames, caller expects %d", path, r.n_frames(), n_frames);
        return XRB_ERR_INVALID;
    }
    int64_t row = 0, nb = 0;
    std::string name;
    row_offsets[0] = 0;
    if (name_offsets) name_offsets[0] = 0;
    for (int i = 0; i < n_frames; ++i) {
        int32_t np = 0;
        if ((rc = r.header(&name, &np))) return rc;
        if ((rc = r.keypoints(keypoints ? keypoints + 4 * row : nullptr))) return rc;
        if (!desc_block && np) {
            set_error("ftr_read: desc_block is NULL");
            return XRB_ERR_INVALID;
        }
        if ((rc = r.descriptors(np ? desc_block + 128 * row : nullptr))) return rc;
        if (names) memcpy(names + nb, name.c_str(), name.size() + 1);
        row += np, nb += (int64_t)name.size() + 1;
        row_offsets[i + 1] = row;
        if (name_offsets) name_offsets[i + 1] = nb;
    }
    return XRB_OK;
}

int xrb_ftr_write(const char *path, int32_t n_frames, const int64_t *row_offsets, const uint8_t *desc_block,
                  const float *keypoints, const char *names, const int64_t *name_offsets) {
    if (!path || n_frames < 0 || !row_offsets || (names && !name_offsets) ||
        (n_frames && row_offsets[n_frames] > row_offsets[0] && !desc_block)) {
        set_error("ftr_write: bad arguments");
        return XRB_ERR_INVALID;
    }
    BinFile f;
    if (!f.open_write(path)) return cannot_open(path);
    bool ok = f.put(n_frames);
    const std::vector<float> zeros(4 * 4096, 0.0f);
    for (int i = 0; i < n_frames && ok; ++i) {
        const int64_t r0 = row_offsets[i], np64 = row_offsets[i + 1] - r0;
        if (np64 < 0 || np64 > 1000000) {
            set_error("ftr_write: frame %d has %lld points (the reader accepts at most 1e6)", i, (long long)np64);
            return XRB_ERR_INVALID;
        }
        const int32_t np = (int32_t)np64;
        if (names) {
            const char *nm = names + name_offsets[i];
            ok = ok && f.write(nm, strlen(nm) + 1);  // write_name: string + NUL (io_base.hpp:84-87)
        } else {
            ok = ok && f.put((char)0);
        }
        ok = ok && f.put(np);
        if (keypoints) {
            ok = ok && f.write(keypoints + 4 * r0, (size_t)np * 16);
        } else {
            for (int32_t k = 0; k < np && ok; k += 4096) ok = f.write(zeros.data(), (size_t)std::min(4096, np - k) * 16);
        }
        ok = ok && f.write(desc_block + 128 * r0, (size_t)np * 128);
    }
    ok = f.close() && ok;
    if (!ok) {
        set_error("ftr_write: write to %s failed", path);
        return XRB_ERR_INVALID;
    }
    return XRB_OK;
}

// ---- fp.bin --------------------------------------------------------------------------------

int xrb_fp_scan(const char *path, int64_t *n_pairs, int64_t *total_matches) {
    if (!path || !n_pairs || !total_matches) {
        set_error("fp_scan: bad arguments");
        return XRB_ERR_INVALID;
    }
    int64_t np = 0, tm = 0;
    const int rc = fp_walk(path, [&](const FpPairHeader &h, BinFile &f) -> int {
        if (!f.skip((int64_t)h.n_matches * 16 + 72 + 4 + (int64_t)h.n_matches)) return bad_file(f, "fp.bin: truncated pair");
        if (h.id1 != h.id2) ++np, tm += (int64_t)h.n_matches;  // io_feature.hpp:120-126
        return XRB_OK;
    });
    if (rc) return rc;
    *n_pairs = np, *total_matches = tm;
    return XRB_OK;
}

int xrb_fp_read(const char *path, int64_t n_pairs, int32_t (*ids)[2], int64_t *offsets, int32_t (*matches)[2],
                double *distances, double *E, int32_t *inlier_num, char *inlier_mask) {
    if (!path || n_pairs < 0 || !offsets || (n_pairs && !ids)) {
        set_error("fp_read: bad arguments");
        return XRB_ERR_INVALID;
    }
    int64_t p = 0, m0 = 0;
    offsets[0] = 0;
    std::vector<unsigned char> buf;
    const int rc = fp_walk(path, [&](const FpPairHeader &h, BinFile &f) -> int {
        const int64_t nm = (int64_t)h.n_matches;
        if (h.id1 == h.id2) {
            if (!f.skip(nm * 16 + 72 + 4 + nm)) return bad_file(f, "fp.bin: truncated pair");
            return XRB_OK;
        }
        if (p >= n_pairs) {
            set_error("%s holds more than the %lld pairs the caller expects", f.path(), (long long)n_pairs);
            return XRB_ERR_INVALID;
        }
        if (nm && !matches) {
            set_error("fp_read: matches is NULL");
            return XRB_ERR_INVALID;
        }
        buf.resize((size_t)nm * 16);
        if (!f.read(buf.data(), buf.size())) return bad_file(f, "fp.bin: truncated matches");
        for (int64_t k = 0; k < nm; ++k) {  // Match{int id1; int id2; double distance;} (types.h:14-21)
            memcpy(matches[m0 + k], buf.data() + 16 * k, 8);
            if (distances) memcpy(&distances[m0 + k], buf.data() + 16 * k + 8, 8);
        }
        double e[9];
        int32_t inl = 0;
        if (!f.read(e, 72) || !f.get(&inl)) return bad_file(f, "fp.bin: truncated pair tail");
        if (E) memcpy(E + 9 * p, e, 72);
        if (inlier_num) inlier_num[p] = inl;
        if (!(inlier_mask ? f.read(inlier_mask + m0, (size_t)nm) : f.skip(nm))) return bad_file(f, "fp.bin: truncated inlier mask");
        ids[p][0] = h.id1, ids[p][1] = h.id2;
        m0 += nm, ++p;
        offsets[p] = m0;
        return XRB_OK;
    });
    if (rc) return rc;
    if (p != n_pairs) {
        set_error("%s holds %lld pairs (self-pairs dropped), caller expects %lld", path, (long long)p, (long long)n_pairs);
        return XRB_ERR_INVALID;
    }
    return XRB_OK;
}

int xrb_fp_write(const char *path, int64_t n_pairs, const int32_t (*ids)[2], const int64_t *offsets,
                 const int32_t (*matches)[2], const double *distances, const double *E, const int32_t *inlier_num,
                 const char *inlier_mask) {
    if (!path || n_pairs < 0 || !offsets || (n_pairs && !ids) || (n_pairs && offsets[n_pairs] > offsets[0] && !matches)) {
        set_error("fp_write: bad arguments");
        return XRB_ERR_INVALID;
    }
    BinFile f;
    if (!f.open_write(path)) return cannot_open(path);
    const uint64_t n = (uint64_t)n_pairs;  // size_t in the reference (io_feature.hpp:134-135)
    bool ok = f.put(n);
    std::vector<unsigned char> buf;
    std::vector<char> ones;
    const double zeroE[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t p = 0; p < n_pairs && ok; ++p) {
        const int64_t m0 = offsets[p], nm = offsets[p + 1] - m0;
        if (nm < 0) {
            set_error("fp_write: offsets decrease at pair %lld", (long long)p);
            return XRB_ERR_INVALID;
        }
        const uint64_t nm64 = (uint64_t)nm;
        ok = ok && f.put(ids[p][0]) && f.put(ids[p][1]) && f.put(nm64);
        buf.resize((size_t)nm * 16);
        for (int64_t k = 0; k < nm; ++k) {
            const double d = distances ? distances[m0 + k] : 0.0;
            memcpy(buf.data() + 16 * k, matches[m0 + k], 8);
            memcpy(buf.data() + 16 * k + 8, &d, 8);
        }
        ok = ok && f.write(buf.data(), buf.size());
        ok = ok && f.write(E ? E + 9 * p : zeroE, 72);
        int32_t inl = 0;
        if (inlier_num) {
            inl = inlier_num[p];
        } else if (inlier_mask) {
            for (int64_t k = 0; k < nm; ++k) inl += inlier_mask[m0 + k] != 0;
        } else {
            inl = (int32_t)nm;
        }
        ok = ok && f.put(inl);
        if (inlier_mask) {
            ok = ok && f.write(inlier_mask + m0, (size_t)nm);
        } else {
            ones.assign((size_t)nm, (char)1);
            ok = ok && f.write(ones.data(), ones.size());
        }
    }
    ok = f.close() && ok;
    if (!ok) {
        set_error("fp_write: write to %s failed", path);
        return XRB_ERR_INVALID;
    }
    return XRB_OK;
}

// ---- COLMAP-style model --------------------------------------------------------------------

int xrb_colmap_scan(const char *dir, xrb_colmap_sizes *sizes) {
    if (!dir || !sizes) {
        set_error("colmap_scan: bad arguments");
        return XRB_ERR_INVALID;
    }
    memset(sizes, 0, sizeof(*sizes));
    {   // cameras.bin (io_ecim.cc:9-29)
        BinFile f;
        const std::string path = join(dir, "cameras.bin");
        if (!f.open_read(path.c_str())) return cannot_open(path.c_str());
        uint64_t n = 0;
        if (!f.get(&n) || n > (uint64_t)f.size() / 24) return bad_file(f, "cameras.bin: bad camera count");
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t id, model;
            if (!f.get(&id) || !f.get(&model) || model > 4) return bad_file(f, "cameras.bin: unknown camera model id");
            if (!f.skip(16 + 8 * kCamParams[model])) return bad_file(f, "cameras.bin: truncated camera");
        }
        sizes->n_cameras = (int32_t)n;
    }
    std::unordered_map<uint64_t, int32_t> track_index;
    int64_t n_points = 0, n_frames = 0, n_p2d = 0, n_obs = 0;
    int rc = walk_points(join(dir, "points3D.bin"), [&](int64_t i, uint64_t id, const double *) { track_index[id] = (int32_t)i; },
                         &n_points);
    if (rc) return rc;
    rc = walk_images(join(dir, "images.bin"), [&](int64_t, const FrameHeader &) { return XRB_OK; },
                     [&](int64_t, int64_t, double, double, uint64_t tid) {
                         ++n_p2d;
                         if (tid != ~0ull && track_index.count(tid)) ++n_obs;  // track_ids_[i] == -1 skipped (ba_solver.cc:337)
                     },
                     &n_frames);
    if (rc) return rc;
    if (n_points > INT32_MAX || n_frames > INT32_MAX || n_obs > INT32_MAX) {
        set_error("colmap_scan: model too large for 32-bit indices");
        return XRB_ERR_INVALID;
    }
    sizes->n_points = (int32_t)n_points, sizes->n_frames = (int32_t)n_frames;
    sizes->n_p2d = n_p2d, sizes->n_obs = n_obs;
    return XRB_OK;
}

int xrb_colmap_read_problem(const char *dir, const xrb_colmap_sizes *sizes, xrb_ba_problem *prob, int32_t *frame_ids,
                            int32_t *camera_ids, uint64_t *track_ids, int32_t *obs_p2d) {
    if (!dir || !sizes || !prob) {
        set_error("colmap_read_problem: bad arguments");
        return XRB_ERR_INVALID;
    }
    if (prob->n_cams != sizes->n_frames || prob->n_pts != sizes->n_points || prob->n_obs != sizes->n_obs ||
        prob->n_intr != sizes->n_cameras || !prob->cam_q || !prob->cam_t || !prob->pts || !prob->intr ||
        !prob->intr_model || !prob->cam_intr || (prob->n_obs && (!prob->obs_cam || !prob->obs_pt || !prob->obs_uv))) {
        set_error("colmap_read_problem: xrb_ba_problem not allocated for the scanned sizes");
        return XRB_ERR_INVALID;
    }
    double *intr = const_cast<double *>(prob->intr);
    int32_t *intr_model = const_cast<int32_t *>(prob->intr_model), *cam_intr = const_cast<int32_t *>(prob->cam_intr);
    int32_t *obs_cam = const_cast<int32_t *>(prob->obs_cam), *obs_pt = const_cast<int32_t *>(prob->obs_pt);
    double *obs_uv = const_cast<double *>(prob->obs_uv);
    std::unordered_map<uint32_t, int32_t> cam_index;
    {
        BinFile f;
        const std::string path = join(dir, "cameras.bin");
        if (!f.open_read(path.c_str())) return cannot_open(path.c_str());
        uint64_t n = 0;
        if (!f.get(&n) || (int64_t)n != sizes->n_cameras) return bad_file(f, "cameras.bin: camera count changed since the scan");
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t id, model;
            uint64_t wh[2];
            if (!f.get(&id) || !f.get(&model) || model > 4 || !f.read(wh, 16)) return bad_file(f, "cameras.bin: bad camera");
            double *p = intr + 8 * i;
            for (int k = 0; k < 8; ++k) p[k] = 0.0;
            if (!f.read(p, 8 * (size_t)kCamParams[model])) return bad_file(f, "cameras.bin: truncated parameters");
            intr_model[i] = (int32_t)model;
            cam_index[id] = (int32_t)i;
            if (camera_ids) camera_ids[i] = (int32_t)id;
        }
    }
    std::unordered_map<uint64_t, int32_t> track_index;
    int64_t n_points = 0, n_frames = 0, o = 0;
    int rc = walk_points(join(dir, "points3D.bin"),
                         [&](int64_t i, uint64_t id, const double *xyz) {
                             if (i < sizes->n_points) {
                                 memcpy(prob->pts + 3 * i, xyz, 24);
                                 if (track_ids) track_ids[i] = id;
                             }
                             track_index[id] = (int32_t)i;
                         },
                         &n_points);
    if (rc) return rc;
    if (n_points != sizes->n_points) {
        set_error("points3D.bin changed since the scan");
        return XRB_ERR_INVALID;
    }
    bool overflow = false;
    rc = walk_images(join(dir, "images.bin"),
                     [&](int64_t i, const FrameHeader &h) -> int {
                         if (i >= sizes->n_frames) {
                             set_error("images.bin changed since the scan");
                             return XRB_ERR_INVALID;
                         }
                         const auto it = cam_index.find(h.camera_id);
                         if (it == cam_index.end()) {
                             set_error("images.bin: frame %u uses camera %u, which cameras.bin does not hold", h.id, h.camera_id);
                             return XRB_ERR_INVALID;
                         }
                         double *q = prob->cam_q + 4 * i;
                         q[0] = h.q[1], q[1] = h.q[2], q[2] = h.q[3], q[3] = h.q[0];  // file: w x y z -> Eigen coeffs x y z w
                         memcpy(prob->cam_t + 3 * i, h.t, 24);
                         cam_intr[i] = it->second;
                         if (frame_ids) frame_ids[i] = (int32_t)h.id;
                         return XRB_OK;
                     },
                     [&](int64_t i, int64_t k, double x, double y, uint64_t tid) {
                         if (tid == ~0ull) return;
                         const auto it = track_index.find(tid);
                         if (it == track_index.end()) return;
                         if (o >= sizes->n_obs) {
                             overflow = true;
                             return;
                         }
                         obs_cam[o] = (int32_t)i, obs_pt[o] = it->second;
                         obs_uv[2 * o] = x, obs_uv[2 * o + 1] = y;
                         if (obs_p2d) obs_p2d[o] = (int32_t)k;
                         ++o;
                     },
                     &n_frames);
    if (rc) return rc;
    if (overflow || o != sizes->n_obs || n_frames != sizes->n_frames) {
        set_error("images.bin changed since the scan");
        return XRB_ERR_INVALID;
    }
    return XRB_OK;
}

int xrb_colmap_write_updated(const char *dir_in, const char *dir_out, const xrb_colmap_sizes *sizes,
                             const xrb_ba_problem *prob) {
    if (!dir_in || !dir_out || !sizes || !prob || prob->n_cams != sizes->n_frames || prob->n_pts != sizes->n_points ||
        !prob->cam_q || !prob->cam_t || !prob->pts) {
        set_error("colmap_write_updated: bad arguments");
        return XRB_ERR_INVALID;
    }
    std::vector<unsigned char> buf;
    // WriteColMapDataBinary creates the directory (io_ecim.cc); every file is written next to its target and
    // renamed into place once complete, so dir_out == dir_in (write back after BA) is safe
    make_dirs(dir_out);
    auto publish = [](BinFile &in, BinFile &out, const std::string &tmp, const std::string &dst, bool ok) {
        in.close();
        ok = out.close() && ok;
        if (ok && rename(tmp.c_str(), dst.c_str()) != 0) ok = false;
        if (!ok) remove(tmp.c_str());
        return ok;
    };
    {   // cameras.bin: intrinsics are constant in the reference's BA (ba_solver.cc:608) — plain copy
        BinFile in, out;
        const std::string pi = join(dir_in, "cameras.bin"), po = join(dir_out, "cameras.bin"), pt = po + ".xrb_tmp";
        if (!in.open_read(pi.c_str())) return cannot_open(pi.c_str());
        if (!out.open_write(pt.c_str())) return cannot_open(pt.c_str());
        buf.resize((size_t)in.size());
        const bool ok = in.read(buf.data(), buf.size()) && out.write(buf.data(), buf.size());
        if (!publish(in, out, pt, po, ok)) return bad_file(out, "cameras.bin: copy failed");
    }
    {   // images.bin: q (w x y z) and t replaced frame by frame
        BinFile in, out;
        const std::string pi = join(dir_in, "images.bin"), po = join(dir_out, "images.bin"), pt = po + ".xrb_tmp";
        if (!in.open_read(pi.c_str())) return cannot_open(pi.c_str());
        if (!out.open_write(pt.c_str())) return cannot_open(pt.c_str());
        uint64_t n = 0;
        if (!in.get(&n) || (int64_t)n != sizes->n_frames) return bad_file(in, "images.bin: frame count differs from sizes");
        bool ok = out.put(n);
        for (uint64_t i = 0; i < n && ok; ++i) {
            uint32_t id, cam;
            double qt[7];
            std::string name;
            uint64_t n_p2d;
            if (!in.get(&id) || !in.read(qt, 56) || !in.get(&cam) || !in.read_name(&name) || !in.get(&n_p2d) ||
                n_p2d > (uint64_t)in.size() / 24)
                return bad_file(in, "images.bin: truncated frame header");
            const double *q = prob->cam_q + 4 * i, *t = prob->cam_t + 3 * i;
            const double qw[4] = {q[3], q[0], q[1], q[2]};
            buf.resize((size_t)n_p2d * 24);
            if (!in.read(buf.data(), buf.size())) return bad_file(in, "images.bin: truncated 2-D points");
            ok = out.put(id) && out.write(qw, 32) && out.write(t, 24) && out.put(cam) &&
                 out.write(name.c_str(), name.size() + 1) && out.put(n_p2d) && out.write(buf.data(), buf.size());
        }
        if (!publish(in, out, pt, po, ok)) return bad_file(out, "images.bin: write failed");
    }
    {   // points3D.bin: xyz replaced track by track
        BinFile in, out;
        const std::string pi = join(dir_in, "points3D.bin"), po = join(dir_out, "points3D.bin"), pt = po + ".xrb_tmp";
        if (!in.open_read(pi.c_str())) return cannot_open(pi.c_str());
        if (!out.open_write(pt.c_str())) return cannot_open(pt.c_str());
        uint64_t n = 0;
        if (!in.get(&n) || (int64_t)n != sizes->n_points) return bad_file(in, "points3D.bin: track count differs from sizes");
        bool ok = out.put(n);
        for (uint64_t i = 0; i < n && ok; ++i) {
            uint64_t id, n_obs;
            double xyz[3], err;
            unsigned char rgb[3];
            if (!in.get(&id) || !in.read(xyz, 24) || !in.read(rgb, 3) || !in.get(&err) || !in.get(&n_obs) ||
                n_obs > (uint64_t)in.size() / 8)
                return bad_file(in, "points3D.bin: truncated track");
            buf.resize((size_t)n_obs * 8);
            if (!in.read(buf.data(), buf.size())) return bad_file(in, "points3D.bin: truncated observations");
            ok = out.put(id) && out.write(prob->pts + 3 * i, 24) && out.write(rgb, 3) && out.put(err) && out.put(n_obs) &&
                 out.write(buf.data(), buf.size());
        }
        if (!publish(in, out, pt, po, ok)) return bad_file(out, "points3D.bin: write failed");
    }
    return XRB_OK;
}

}  // extern "C"
