// Host side, no CUDA: scans in, poses out — what the reference's DataIo does on the two
// sides of the hot path, reading straight into the pcl::PointXYZINormal rows (48 bytes: x y z 1 | normal_x normal_y
// normal_z 0 | intensity curvature 0 0) that every registration / front-end entry point of the C-ABI takes, so that a
// caller's pinned buffer (mulls_host_alloc) is filled once and shipped as it is.
//   read_scan        <- DataIo::read_pc_cloud_block   include/common/dataio.hpp:1732-1756
//                         read_pcd_file  :279-287 (pcl::io::loadPCDFile: PCD v0.7, DATA ascii | binary, float32 fields)
//                         read_bin_file  :357-377 (KITTI velodyne: x y z reflectance, intensity = reflectance * 255;
//                                                  the read loop tests eof() only after the failed read, so ONE
//                                                  default-constructed point follows the data — reproduced)
//   append_pose      <- DataIo::write_lo_pose_overwrite / write_lo_pose_append   :1896-1926 (setprecision(8), 12 values)
#pragma once
#include <cctype>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace mulls_io {

enum { kOk = 0, kArg = -1, kCapacity = -2, kUnsupported = -6, kIo = -7 };

inline bool ends_with_ci(const char *s, const char *suffix) {
    const size_t n = std::strlen(s), m = std::strlen(suffix);
    if (m > n) return false;
    for (size_t i = 0; i < m; ++i)
        if (std::tolower((unsigned char)s[n - m + i]) != std::tolower((unsigned char)suffix[i])) return false;
    return true;
}

struct PcdHeader {
    std::vector<std::string> fields;
    std::vector<int> sizes, counts;
    std::vector<char> types;
    size_t points = 0;
    bool have_points = false;
    int data = -1; // 0 ascii, 1 binary, 2 binary_compressed
    long data_offset = 0;
};

// column of the 12-float row a PCD field lands in (-1: ignored)
inline int field_column(const std::string &name) {
    static const char *names[] = {"x", "y", "z", "normal_x", "normal_y", "normal_z", "intensity", "curvature"};
    static const int cols[] = {0, 1, 2, 4, 5, 6, 8, 9};
    for (int i = 0; i < 8; ++i)
        if (name == names[i]) return cols[i];
    return -1;
}

inline int read_pcd_header(FILE *f, PcdHeader &h) {
    char line[4096];
    while (std::fgets(line, sizeof(line), f)) {
        std::vector<std::string> tok;
        for (char *p = std::strtok(line, " \t\r\n"); p; p = std::strtok(nullptr, " \t\r\n")) tok.push_back(p);
        if (tok.empty() || tok[0][0] == '#') continue;
        std::string key = tok[0];
        for (char &c : key) c = (char)std::toupper((unsigned char)c);
        if (key == "FIELDS") h.fields.assign(tok.begin() + 1, tok.end());
        else if (key == "SIZE")
            for (size_t i = 1; i < tok.size(); ++i) h.sizes.push_back(std::atoi(tok[i].c_str()));
        else if (key == "TYPE")
            for (size_t i = 1; i < tok.size(); ++i) h.types.push_back(tok[i][0]);
        else if (key == "COUNT")
            for (size_t i = 1; i < tok.size(); ++i) h.counts.push_back(std::atoi(tok[i].c_str()));
        else if (key == "POINTS" && tok.size() > 1) h.points = (size_t)std::strtoull(tok[1].c_str(), nullptr, 10), h.have_points = true;
        else if (key == "WIDTH" && tok.size() > 1 && !h.have_points) h.points = (size_t)std::strtoull(tok[1].c_str(), nullptr, 10);
        else if (key == "DATA" && tok.size() > 1) {
            std::string d = tok[1];
            for (char &c : d) c = (char)std::tolower((unsigned char)c);
            h.data = d == "ascii" ? 0 : (d == "binary" ? 1 : 2);
            h.data_offset = std::ftell(f);
            break;
        }
    }
    if (h.fields.empty() || h.data < 0) return kIo;
    if (h.counts.empty()) h.counts.assign(h.fields.size(), 1);
    if (h.sizes.size() != h.fields.size() || h.types.size() != h.fields.size() || h.counts.size() != h.fields.size()) return kIo;
    for (size_t i = 0; i < h.fields.size(); ++i)
        if (h.sizes[i] != 4 || (h.types[i] != 'F' && h.types[i] != 'f') || h.counts[i] != 1) return kUnsupported;
    if (h.data == 2) return kUnsupported; // binary_compressed (LZF) is not handled
    return kOk;
}

inline void default_row(float *r) {
    for (int k = 0; k < 12; ++k) r[k] = 0.0f;
    r[3] = 1.0f; // pcl::PointXYZINormal(): data[3] = 1
}

// number of rows read_scan will produce for `path`
inline int probe_scan(const char *path, size_t *n_points) {
    if (!path || !n_points) return kArg;
    FILE *f = std::fopen(path, "rb");
    if (!f) return kIo;
    int rc = kOk;
    if (ends_with_ci(path, ".bin")) {
        std::fseek(f, 0, SEEK_END);
        const long bytes = std::ftell(f);
        *n_points = (size_t)(bytes / 16) + 1; // + the reference's end-of-file point
    } else {
        PcdHeader h;
        rc = read_pcd_header(f, h);
        if (rc == kOk) *n_points = h.points;
    }
    std::fclose(f);
    return rc;
}

inline int read_scan(const char *path, float *rows, size_t cap, size_t *n_points, double bound[6], int normalize_intensity) {
    if (!path || !rows || !n_points) return kArg;
    FILE *f = std::fopen(path, "rb");
    if (!f) return kIo;
    size_t n = 0;
    int rc = kOk;
    if (ends_with_ci(path, ".bin")) {
        std::fseek(f, 0, SEEK_END);
        const size_t full = (size_t)(std::ftell(f) / 16);
        std::fseek(f, 0, SEEK_SET);
        n = full + 1;
        if (n > cap) rc = kCapacity;
        else {
            float rec[4];
            for (size_t i = 0; i < full && rc == kOk; ++i) {
                if (std::fread(rec, 4, 4, f) != 4) {
                    rc = kIo;
                    break;
                }
                float *r = rows + 12 * i;
                default_row(r);
                r[0] = rec[0], r[1] = rec[1], r[2] = rec[2];
                r[8] = rec[3] * 255.0f; // point.intensity *= 255
            }
            default_row(rows + 12 * full); // the read that hit end-of-file left the point default-constructed
        }
    } else {
        PcdHeader h;
        rc = read_pcd_header(f, h);
        if (rc == kOk) {
            n = h.points;
            const size_t nf = h.fields.size();
            std::vector<int> col(nf);
            for (size_t j = 0; j < nf; ++j) col[j] = field_column(h.fields[j]);
            if (n > cap) rc = kCapacity;
            else if (h.data == 1) {
                std::vector<float> rec(nf);
                for (size_t i = 0; i < n && rc == kOk; ++i) {
                    if (std::fread(rec.data(), 4, nf, f) != nf) {
                        rc = kIo;
                        break;
                    }
                    float *r = rows + 12 * i;
                    default_row(r);
                    for (size_t j = 0; j < nf; ++j)
                        if (col[j] >= 0) r[col[j]] = rec[j];
                }
            } else {
                char line[8192];
                size_t i = 0;
                while (i < n && std::fgets(line, sizeof(line), f)) {
                    char *p = line;
                    while (*p && std::isspace((unsigned char)*p)) ++p;
                    if (!*p) continue;
                    float *r = rows + 12 * i;
                    default_row(r);
                    for (size_t j = 0; j < nf; ++j) {
                        char *end = nullptr;
                        const float v = std::strtof(p, &end); // (decimal -> float directly, as the reference's stream extraction)
                        if (end == p) {
                            rc = kIo;
                            break;
                        }
                        if (col[j] >= 0) r[col[j]] = v;
                        p = end;
                    }
                    if (rc != kOk) break;
                    ++i;
                }
                if (rc == kOk && i < n) rc = kIo;
            }
        }
    }
    std::fclose(f);
    if (rc != kOk) return rc;
    *n_points = n;
    // CloudUtility::get_cloud_bbx (utility.hpp:817-848)
    if (bound) {
        double b[6] = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (size_t i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) {
                const double v = rows[12 * i + d];
                if (b[d] > v) b[d] = v;
                if (b[3 + d] < v) b[3 + d] = v;
            }
        for (int d = 0; d < 6; ++d) bound[d] = b[d];
    }
    // dataio.hpp:1738-1750: intensity rescaled to 0..255 with float arithmetic
    if (normalize_intensity && n) {
        float lo = FLT_MAX, hi = -FLT_MAX;
        for (size_t i = 0; i < n; ++i) {
            const float v = rows[12 * i + 8];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        const float scale = (float)(255.0 / (double)(hi - lo)); // float intesnity_scale = 255.0 / (max - min)
        for (size_t i = 0; i < n; ++i) rows[12 * i + 8] = (rows[12 * i + 8] - lo) * scale;
    }
    return kOk;
}

inline int append_pose(const char *path, const double T[16], int overwrite) {
    if (!path || !T) return kArg;
    FILE *f = std::fopen(path, overwrite ? "w" : "a");
    if (!f) return kIo;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) std::fprintf(f, "%s%.8g", (r || c) ? " " : "", T[4 * r + c]);
    std::fprintf(f, "\n");
    std::fclose(f);
    return kOk;
}

} // namespace mulls_io
