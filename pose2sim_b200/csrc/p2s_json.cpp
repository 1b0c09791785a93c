// Host staging: OpenPose-format JSON files -> packed float32 observation planes, multi-threaded.
//
// Replaces the per-frame / per-person file handling of the reference
//   Pose2Sim/triangulation.py:607-653  extract_files_frame_f   (re-opens every file once per person)
//   Pose2Sim/triangulation.py:77-90    count_persons_in_json
// with ONE parse per file.  The value taken for (person n, keypoint id) is what
// `js['people'][n]['pose_keypoints_2d'][3*id : 3*id+3]` yields in Python; anything that would raise
// there (missing / unparsable file, person or key absent, list too short, non-numeric entry) is a NaN
// triple, like the reference's bare `except:` (:629-644).
//
// The parser validates the WHOLE document with the grammar Python's `json.load` accepts (RFC 8259 plus
// the NaN / Infinity / -Infinity literals; duplicate keys: the last one wins), because a syntax error
// anywhere makes `json.load` raise and the reference then treats the file as missing.
#include <atomic>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <algorithm>
#include <dirent.h>
#include <fcntl.h>
#include <unistd.h>
#include <new>
#include <sys/stat.h>

#include "../../include/pose2sim_b200.h"

namespace {

struct Person {
    bool is_object = false;
    bool has_keypoints = false;      // key present and its value is an array
    bool bad_entry = false;          // an array element that float() would reject (string, array, object)
    bool odd_entry = false;          // an element that is not a plain JSON number (null, true, false, NaN, Infinity)
    std::vector<double> kp;
};

struct Doc {
    bool has_people = false;         // top level is an object with a "people" array
    std::vector<Person> people;
};

class Parser {
  public:
    Parser(const char *b, const char *e) : p_(b), end_(e) {}

    bool parse(Doc &doc) {
        ws();
        if (!top(doc)) return false;
        ws();
        return p_ == end_;           // trailing data is an error in json.load
    }

  private:
    const char *p_, *end_;
    int depth_ = 0;

    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool lit(const char *s) {
        const size_t n = std::strlen(s);
        if ((size_t)(end_ - p_) < n || std::memcmp(p_, s, n) != 0) return false;
        p_ += n;
        return true;
    }

    bool string(std::string *out) {
        if (p_ >= end_ || *p_ != '"') return false;
        ++p_;
        while (p_ < end_) {
            const unsigned char c = (unsigned char)*p_;
            if (c == '"') { ++p_; return true; }
            if (c < 0x20) return false;                          // control characters are rejected (strict mode)
            if (c == '\\') {
                if (++p_ >= end_) return false;
                const char e = *p_;
                if (e == 'u') {
                    if (end_ - p_ < 5) return false;
                    for (int i = 1; i <= 4; ++i) if (!std::isxdigit((unsigned char)p_[i])) return false;
                    if (out) out->push_back('?');                // keys of interest are ASCII
                    p_ += 5;
                    continue;
                }
                if (!std::strchr("\"\\/bfnrt", e)) return false;
                if (out) out->push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
                ++p_;
                continue;
            }
            if (out) out->push_back((char)c);
            ++p_;
        }
        return false;
    }

    // JSON number grammar: -?(0|[1-9]\d*)(\.\d+)?([eE][-+]?\d+)?
    bool number(double *out) {
        const char *s = p_;
        if (p_ < end_ && *p_ == '-') ++p_;
        if (p_ >= end_) return false;
        if (*p_ == '0') ++p_;
        else if (*p_ >= '1' && *p_ <= '9') { while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
        else return false;
        if (p_ < end_ && *p_ == '.') {
            ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            ++p_;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (out) {
            double v = 0.0;
            auto r = std::from_chars(s, p_, v);                  // correctly rounded, locale independent
            if (r.ec == std::errc::result_out_of_range) {        // overflow -> +-inf, underflow -> +-0 / denormal, like float()
                const std::string tok(s, p_);
                v = std::strtod(tok.c_str(), nullptr);
            }
            else if (r.ec != std::errc()) return false;
            *out = v;
        }
        return true;
    }

    // scalar-or-container; when `num` is given and the value is numeric-like, *num receives what
    // numpy's float conversion would give (true -> 1, false -> 0, null -> NaN) and *is_num = true
    bool value(double *num, bool *is_num, bool *is_plain = nullptr) {
        if (is_num) *is_num = false;
        if (is_plain) *is_plain = false;
        ws();
        if (p_ >= end_) return false;
        const char c = *p_;
        if (c == '{') return object(nullptr, nullptr);
        if (c == '[') return array_skip();
        if (c == '"') return string(nullptr);
        double v;
        bool ok;
        if (c == 't') { ok = lit("true"); v = 1.0; }
        else if (c == 'f') { ok = lit("false"); v = 0.0; }
        else if (c == 'n') { ok = lit("null"); v = std::numeric_limits<double>::quiet_NaN(); }
        else if (c == 'N') { ok = lit("NaN"); v = std::numeric_limits<double>::quiet_NaN(); }
        else if (c == 'I') { ok = lit("Infinity"); v = HUGE_VAL; }
        else if (c == '-' && p_ + 1 < end_ && p_[1] == 'I') { ok = lit("-Infinity"); v = -HUGE_VAL; }
        else { ok = number(&v); if (is_plain) *is_plain = ok; }
        if (!ok) return false;
        if (num) *num = v;
        if (is_num) *is_num = true;
        return true;
    }

    bool array_skip() {
        if (++depth_ > 512) return false;
        ++p_;
        ws();
        if (p_ < end_ && *p_ == ']') { ++p_; --depth_; return true; }
        for (;;) {
            if (!value(nullptr, nullptr)) return false;
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; --depth_; return true; }
            return false;
        }
    }

    bool number_array(Person &person) {
        person.kp.clear();
        person.bad_entry = false;
        person.odd_entry = false;
        ++p_;
        ws();
        if (p_ < end_ && *p_ == ']') { ++p_; return true; }
        for (;;) {
            double v;
            bool is_num, is_plain;
            if (!value(&v, &is_num, &is_plain)) return false;
            if (!is_plain) person.odd_entry = true;
            if (is_num) person.kp.push_back(v);
            else { person.bad_entry = true; person.kp.push_back(std::numeric_limits<double>::quiet_NaN()); }
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; return true; }
            return false;
        }
    }

    // generic object; `person` != null: capture "pose_keypoints_2d"; `doc` != null: capture "people"
    bool object(Person *person, Doc *doc) {
        if (++depth_ > 512) return false;
        ++p_;
        ws();
        if (p_ < end_ && *p_ == '}') { ++p_; --depth_; return true; }
        std::string key;
        for (;;) {
            ws();
            key.clear();
            if (!string(&key)) return false;
            ws();
            if (p_ >= end_ || *p_ != ':') return false;
            ++p_;
            ws();
            if (person && key == "pose_keypoints_2d") {
                if (p_ < end_ && *p_ == '[') {
                    person->has_keypoints = true;
                    if (!number_array(*person)) return false;
                } else {                                         // not a list: indexing it raises in Python
                    person->has_keypoints = false;
                    person->kp.clear();
                    if (!value(nullptr, nullptr)) return false;
                }
            } else if (doc && key == "people") {
                doc->people.clear();
                if (p_ < end_ && *p_ == '[') {
                    doc->has_people = true;
                    if (!people_array(*doc)) return false;
                } else {
                    doc->has_people = false;
                    if (!value(nullptr, nullptr)) return false;
                }
            } else if (!value(nullptr, nullptr)) {
                return false;
            }
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == '}') { ++p_; --depth_; return true; }
            return false;
        }
    }

    bool people_array(Doc &doc) {
        ++p_;
        ws();
        if (p_ < end_ && *p_ == ']') { ++p_; return true; }
        for (;;) {
            ws();
            doc.people.emplace_back();
            Person &person = doc.people.back();
            if (p_ < end_ && *p_ == '{') {
                person.is_object = true;
                if (!object(&person, nullptr)) return false;
            } else if (!value(nullptr, nullptr)) {
                return false;
            }
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; return true; }
            return false;
        }
    }

    bool top(Doc &doc) {
        if (p_ < end_ && *p_ == '{') return object(nullptr, &doc);
        return value(nullptr, nullptr);
    }
};

bool read_file(const char *path, std::string &buf) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    buf.clear();
    char tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    const bool ok = !std::ferror(f);
    std::fclose(f);
    return ok;
}

}  // namespace

extern "C" int p2s_read_pose_files(const char *const *paths, long long n_frames, int n_cams,
                                   const int32_t *keypoint_ids, int n_keypoints, int n_persons,
                                   float *x, float *y, float *lik, int32_t *n_people, uint8_t *status,
                                   long long *n_inexact, int n_threads) {
    if (!paths || n_frames < 0 || n_cams < 1 || n_keypoints < 0 || n_persons < 0 || (n_keypoints > 0 && !keypoint_ids))
        return P2S_EINVAL;
    if ((long long)n_persons * n_keypoints > 0 && (!x || !y || !lik)) return P2S_EINVAL;
    for (int k = 0; k < n_keypoints; ++k) if (keypoint_ids[k] < 0) return P2S_EINVAL;
    const long long n_files = n_frames * n_cams;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((long long)n_threads > n_files) n_threads = (int)(n_files > 0 ? n_files : 1);
    std::atomic<long long> next{0};
    std::atomic<long long> inexact{0};
    const float fnan = std::numeric_limits<float>::quiet_NaN();
    const size_t C = (size_t)n_cams, K = (size_t)n_keypoints, N = (size_t)n_persons;

    auto work = [&]() {
        std::string buf;
        Doc doc;
        long long my_inexact = 0;
        for (;;) {
            const long long i0 = next.fetch_add(16);
            if (i0 >= n_files) break;
            const long long i1 = i0 + 16 < n_files ? i0 + 16 : n_files;
            for (long long i = i0; i < i1; ++i) {
                const size_t f = (size_t)(i / n_cams), c = (size_t)(i % n_cams);
                doc.has_people = false;
                doc.people.clear();
                bool ok = paths[i] && paths[i][0] && read_file(paths[i], buf);
                if (ok) {
                    Parser ps(buf.data(), buf.data() + buf.size());
                    ok = ps.parse(doc);
                }
                if (status) status[i] = ok ? 1 : 0;
                if (n_people) n_people[i] = (ok && doc.has_people) ? (int32_t)doc.people.size() : (ok ? 0 : -1);
                for (size_t n = 0; n < N; ++n) {
                    const Person *person = (ok && doc.has_people && n < doc.people.size()) ? &doc.people[n] : nullptr;
                    const bool usable = person && person->is_object && person->has_keypoints && !person->bad_entry;
                    for (size_t k = 0; k < K; ++k) {
                        const size_t o = ((f * N + n) * K + k) * C + c;
                        const size_t j = 3 * (size_t)keypoint_ids[k];
                        if (usable && j + 2 < person->kp.size()) {
                            const double vx = person->kp[j], vy = person->kp[j + 1], vl = person->kp[j + 2];
                            const float fx = (float)vx, fy = (float)vy, fl = (float)vl;
                            x[o] = fx; y[o] = fy; lik[o] = fl;
                            my_inexact += (std::isfinite(vx) && (double)fx != vx) + (std::isfinite(vy) && (double)fy != vy) +
                                          (std::isfinite(vl) && (double)fl != vl);
                        } else {
                            x[o] = fnan; y[o] = fnan; lik[o] = fnan;
                        }
                    }
                }
            }
        }
        inexact.fetch_add(my_inexact);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (n_inexact) *n_inexact = inexact.load();
    return P2S_OK;
}

// ---- directory index: listing, the reference's file order and its frame -> file table, natively ---------------------
// What the reference does per stage in Python (triangulation.py:752-803, common.py:568-583): list every camera folder,
// keep *.json, sort by the LAST number in the name (names without a number after those, alphabetically; sorted() is
// stable, so equal numbers keep the listing order), then per frame f pick from every camera ALL its files whose last
// number is f, or 'none', flatten, and use the first n_cams entries (a camera with two files for one frame shifts the
// later cameras).  For a 40 k-file trial those list / regex / dict operations cost more than parsing the files, so they
// live here; staging.py keeps the Python statements as the tested restatement (tests/test_native_staging.py).
struct p2s_dir_index {
    // d0, dn: the last digit run of the name without its leading zeros (a view into `name`, no second string per file)
    struct Entry { std::string name; bool has_num; uint32_t d0, dn; long long num; };
    std::vector<std::string> dirs;
    std::vector<std::vector<Entry>> cams;
    std::string table_arena;                   // the table's paths back to back, NUL-terminated ("" = none)
    std::vector<size_t> table_off;             // [F][C] offsets into the arena
    std::vector<const char *> table;           // [F][C] pointers into the arena
    std::vector<const char *> table_name;      // [F][C] the file name alone (entry of camera c's folder), null = none
};

namespace {

// numeric order of digit strings without leading zeros
bool less_digits(const p2s_dir_index::Entry &a, const p2s_dir_index::Entry &b) {
    return a.dn != b.dn ? a.dn < b.dn : std::memcmp(a.name.data() + a.d0, b.name.data() + b.d0, a.dn) < 0;
}

}  // namespace

extern "C" int p2s_index_open(const char *const *dirs, int n_cams, p2s_dir_index **out) {
    if (!dirs || n_cams < 1 || !out) return P2S_EINVAL;
    *out = nullptr;
    p2s_dir_index *ix = new (std::nothrow) p2s_dir_index();
    if (!ix) return P2S_ENOMEM;
    ix->cams.resize((size_t)n_cams);
    for (int c = 0; c < n_cams; ++c) ix->dirs.emplace_back(dirs[c] ? dirs[c] : "");
    std::vector<int> failed((size_t)n_cams, 0);
    auto list_one = [&](int c) {                                     // one thread per camera folder
        DIR *d = ::opendir(ix->dirs[(size_t)c].c_str());
        if (!d) { failed[(size_t)c] = 1; return; }                   // os.listdir raises: the caller tries the next folder
        std::vector<p2s_dir_index::Entry> &v = ix->cams[(size_t)c];
        while (struct dirent *e = ::readdir(d)) {
            const char *n = e->d_name;
            const size_t len = std::strlen(n);
            if ((len == 1 && n[0] == '.') || (len == 2 && n[0] == '.' && n[1] == '.')) continue;
            if (len < 5 || std::memcmp(n + len - 5, ".json", 5) != 0) continue;     // fnmatch.filter(..., '*.json')
            p2s_dir_index::Entry en;
            en.name.assign(n, len);
            size_t hi = len;
            while (hi > 0 && !(n[hi - 1] >= '0' && n[hi - 1] <= '9')) --hi;        // end of the last digit run
            size_t lo = hi;
            while (lo > 0 && n[lo - 1] >= '0' && n[lo - 1] <= '9') --lo;
            en.has_num = hi > lo;
            en.num = -1;
            en.d0 = en.dn = 0;
            if (en.has_num) {
                size_t z = lo;
                while (z + 1 < hi && n[z] == '0') ++z;
                en.d0 = (uint32_t)z; en.dn = (uint32_t)(hi - z);
                if (en.dn <= 18) {
                    long long v = 0;
                    for (size_t i = z; i < hi; ++i) v = v * 10 + (n[i] - '0');
                    en.num = v;
                }
            }
            v.push_back(std::move(en));
        }
        ::closedir(d);
        std::stable_sort(v.begin(), v.end(), [](const p2s_dir_index::Entry &a, const p2s_dir_index::Entry &b) {
            if (a.has_num != b.has_num) return a.has_num;             // (False, n) < (True, s)
            if (a.has_num) return less_digits(a, b);
            return a.name < b.name;
        });
    };
    {
        std::vector<std::thread> pool;
        for (int c = 1; c < n_cams; ++c) pool.emplace_back(list_one, c);
        list_one(0);
        for (auto &t : pool) t.join();
    }
    for (int c = 0; c < n_cams; ++c) if (failed[(size_t)c]) { delete ix; return P2S_EINVAL; }
    *out = ix;
    return P2S_OK;
}

extern "C" void p2s_index_close(p2s_dir_index *ix) { delete ix; }
extern "C" long long p2s_index_file_count(const p2s_dir_index *ix, int cam) {
    return (ix && cam >= 0 && (size_t)cam < ix->cams.size()) ? (long long)ix->cams[(size_t)cam].size() : -1;
}
extern "C" const char *p2s_index_file_name(const p2s_dir_index *ix, int cam, long long i) {
    if (!ix || cam < 0 || (size_t)cam >= ix->cams.size() || i < 0 || (size_t)i >= ix->cams[(size_t)cam].size()) return nullptr;
    return ix->cams[(size_t)cam][(size_t)i].name.c_str();
}

// frames [f0, f1) with the reference's selection rule; P2S_EINVAL when a listed name has no number (the reference raises
// IndexError at triangulation.py:799 — the caller takes the Python path, which raises it)
extern "C" int p2s_index_build_table(p2s_dir_index *ix, long long f0, long long f1) {
    if (!ix) return P2S_EINVAL;
    const size_t C = ix->cams.size();
    const long long F = f1 > f0 ? f1 - f0 : 0;
    for (const auto &v : ix->cams) for (const auto &e : v) if (!e.has_num) return P2S_EINVAL;
    // per camera: files are sorted by number, so the files of frame f are a contiguous run found by a merge walk
    std::vector<size_t> pos(C, 0);
    ix->table_arena.clear();
    ix->table_off.clear();
    ix->table_name.clear();
    size_t longest = 0;
    for (const auto &v : ix->cams) for (const auto &e : v) longest = std::max(longest, e.name.size());
    size_t dirlen = 0;
    for (const auto &d : ix->dirs) dirlen = std::max(dirlen, d.size());
    ix->table_arena.reserve((size_t)F * C * (dirlen + longest + 2));
    ix->table_off.reserve((size_t)F * C);
    ix->table_name.reserve((size_t)F * C);
    std::vector<const p2s_dir_index::Entry *> flat;
    for (long long f = f0; f < f1; ++f) {
        flat.clear();
        for (size_t c = 0; c < C; ++c) {
            const auto &v = ix->cams[c];
            size_t &p = pos[c];
            if (f == f0) {                                             // first frame: binary search for the run start
                size_t lo = 0, hi = v.size();
                while (lo < hi) { const size_t mid = (lo + hi) / 2; if (v[mid].num >= 0 && v[mid].num < f) lo = mid + 1; else hi = mid; }
                p = lo;
            }
            while (p < v.size() && v[p].num >= 0 && v[p].num < f) ++p;
            size_t q = p;
            bool any = false;
            while (q < v.size() && v[q].num == f) { flat.push_back(&v[q]); ++q; any = true; }
            if (!any) flat.push_back(nullptr);
        }
        // the NAME at flattened position c is joined with camera c's folder (triangulation.py:801-803): after a camera
        // with two files for the frame the later names land in the wrong folder and read as missing, like in the reference
        for (size_t c = 0; c < C; ++c) {
            ix->table_off.push_back(ix->table_arena.size());
            if (flat[c]) {
                ix->table_arena.append(ix->dirs[c]);
                ix->table_arena.push_back('/');
                ix->table_arena.append(flat[c]->name);
                ix->table_name.push_back(flat[c]->name.c_str());
            } else {
                ix->table_name.push_back(nullptr);
            }
            ix->table_arena.push_back('\0');
        }
    }
    ix->table.resize(ix->table_off.size());
    for (size_t i = 0; i < ix->table_off.size(); ++i) ix->table[i] = ix->table_arena.data() + ix->table_off[i];
    return P2S_OK;
}

extern "C" const char *const *p2s_index_table_paths(const p2s_dir_index *ix) { return ix ? ix->table.data() : nullptr; }
extern "C" const char *p2s_index_table_arena(const p2s_dir_index *ix, long long *len) {
    if (len) *len = ix ? (long long)ix->table_arena.size() : 0;
    return ix ? ix->table_arena.data() : nullptr;
}

// ---- file signatures for the staging cache: (mtime in ns, size) per path, -1 / -1 when the file cannot be stat'ed ----
extern "C" int p2s_stat_files(const char *const *paths, long long n, long long *mtime_ns, long long *size, int n_threads) {
    if (!paths || n < 0 || !mtime_ns || !size) return P2S_EINVAL;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((long long)n_threads > n) n_threads = (int)(n > 0 ? n : 1);
    std::atomic<long long> next{0};
    auto work = [&]() {
        for (;;) {
            const long long i0 = next.fetch_add(64);
            if (i0 >= n) break;
            const long long i1 = i0 + 64 < n ? i0 + 64 : n;
            for (long long i = i0; i < i1; ++i) {
                struct stat st;
                if (paths[i] && paths[i][0] && ::stat(paths[i], &st) == 0) {
                    mtime_ns[i] = (long long)st.st_mtim.tv_sec * 1000000000LL + (long long)st.st_mtim.tv_nsec;
                    size[i] = (long long)st.st_size;
                } else {
                    mtime_ns[i] = size[i] = -1;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return P2S_OK;
}

// 128-bit digest of (path, mtime, size) of every table entry: the staging cache key without 40 k Python strings.
// The stat pass runs on all cores with fstatat() relative to each camera folder's descriptor (no walk of the absolute
// path per file); the digest is mixed per 8-byte word.
extern "C" int p2s_index_signature(const p2s_dir_index *ix, unsigned long long sig[2], int n_threads) {
    if (!ix || !sig) return P2S_EINVAL;
    const long long n = (long long)ix->table.size();
    const size_t C = ix->cams.size();
    std::vector<long long> mt((size_t)n), sz((size_t)n);
    if (n > 0) {
        std::vector<int> dfd(C, -1);
        for (size_t c = 0; c < C; ++c) dfd[c] = ::open(ix->dirs[c].c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
        if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads < 1) n_threads = 1;
        if ((long long)n_threads > (n + 255) / 256) n_threads = (int)((n + 255) / 256);
        std::atomic<long long> next{0};
        auto work = [&]() {
            for (;;) {
                const long long i0 = next.fetch_add(256);
                if (i0 >= n) break;
                const long long i1 = i0 + 256 < n ? i0 + 256 : n;
                for (long long i = i0; i < i1; ++i) {
                    struct stat st;
                    const size_t c = (size_t)i % C;
                    const char *nm = ix->table_name[(size_t)i];
                    int rc = -1;
                    if (nm) rc = dfd[c] >= 0 ? ::fstatat(dfd[c], nm, &st, 0) : ::stat(ix->table[(size_t)i], &st);
                    if (rc == 0) {
                        mt[(size_t)i] = (long long)st.st_mtim.tv_sec * 1000000000LL + (long long)st.st_mtim.tv_nsec;
                        sz[(size_t)i] = (long long)st.st_size;
                    } else {
                        mt[(size_t)i] = sz[(size_t)i] = -1;
                    }
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        for (size_t c = 0; c < C; ++c) if (dfd[c] >= 0) ::close(dfd[c]);
    }
    unsigned long long h0 = 1469598103934665603ULL, h1 = 0x9e3779b97f4a7c15ULL;
    auto mix64 = [&](unsigned long long w) {
        h0 = (h0 ^ w) * 1099511628211ULL; h0 ^= h0 >> 29;
        h1 = (h1 + w + (h1 << 6) + (h1 >> 2)) * 0xff51afd7ed558ccdULL; h1 ^= h1 >> 32;
    };
    auto mix = [&](const void *p, size_t len) {
        const unsigned char *b = (const unsigned char *)p;
        size_t i = 0;
        for (; i + 8 <= len; i += 8) { unsigned long long w; std::memcpy(&w, b + i, 8); mix64(w); }
        unsigned long long w = 0x80ULL << 56;
        if (i < len) std::memcpy(&w, b + i, len - i);
        mix64(w ^ (unsigned long long)len);
    };
    for (long long i = 0; i < n; ++i) {
        const char *p = ix->table[(size_t)i];
        mix(p, std::strlen(p) + 1);
        mix64((unsigned long long)mt[(size_t)i]);
        mix64((unsigned long long)sz[(size_t)i]);
    }
    sig[0] = h0; sig[1] = h1;
    return P2S_OK;
}

// ---- association stage: people lists in, rewritten files out -------------------------------------------------------
// Pose2Sim/personAssociation.py reads every camera's JSON once per person COMBINATION (:199-205, :260-274) and rewrites
// the chosen people with json.load + json.dumps per file (:552-580).  Here every source file is parsed once per pass by
// the threads below.  Two index spaces of the reference are kept apart (SURVEY.md 8(a) caveat):
//   A  `persons_combinations` (:81-89): people whose x values are not all NaN  -> `count_named`
//   B  `read_json` (:260-274): keypoint lists with at least 3 values          -> `count_listed`, the order `obs` follows
// Anything irregular — a person that is not an object, a missing or non-list `pose_keypoints_2d`, elements that are
// not plain numbers — gets status 2 and is left to the Python path, which mirrors the reference's exception handling
// statement by statement; regular files (what pose estimators write) never take it.
extern "C" int p2s_read_people_files(const char *const *paths, long long n_frames, int n_cams, int value_offset,
                                     int n_values, int max_persons, float *obs, int32_t *count_named,
                                     int32_t *count_listed, int32_t *list_len, uint8_t *status, long long *n_inexact,
                                     int n_threads) {
    if (!paths || n_frames < 0 || n_cams < 1 || value_offset < 0 || n_values < 0 || max_persons < 0 || !status ||
        !count_named || !count_listed || !list_len || ((long long)n_values * max_persons > 0 && !obs))
        return P2S_EINVAL;
    const long long n_files = n_frames * n_cams;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((long long)n_threads > n_files) n_threads = (int)(n_files > 0 ? n_files : 1);
    std::atomic<long long> next{0}, inexact{0};
    const float fnan = std::numeric_limits<float>::quiet_NaN();
    const size_t L = (size_t)n_values, NP = (size_t)max_persons, off = (size_t)value_offset;
    auto work = [&]() {
        std::string buf;
        Doc doc;
        long long my_inexact = 0;
        for (;;) {
            const long long i0 = next.fetch_add(16);
            if (i0 >= n_files) break;
            const long long i1 = i0 + 16 < n_files ? i0 + 16 : n_files;
            for (long long i = i0; i < i1; ++i) {
                float *o = obs ? obs + (size_t)i * NP * L : nullptr;
                for (size_t j = 0; j < NP * L; ++j) o[j] = fnan;
                count_named[i] = count_listed[i] = list_len[i] = 0;
                doc.has_people = false;
                doc.people.clear();
                bool ok = paths[i] && paths[i][0] && read_file(paths[i], buf);
                if (ok) {
                    Parser ps(buf.data(), buf.data() + buf.size());
                    ok = ps.parse(doc);
                }
                if (!ok) { status[i] = 0; continue; }
                if (!doc.has_people) { status[i] = 2; continue; }          // no `people` list: the Python path decides
                bool regular = true;
                for (const Person &q : doc.people) regular = regular && q.is_object && q.has_keypoints && !q.bad_entry && !q.odd_entry;
                if (!regular) { status[i] = 2; continue; }
                int named = 0, listed = 0, len = 0;
                for (const Person &q : doc.people) {
                    if (!q.kp.empty()) ++named;                             // A: a non-empty x slice of plain numbers
                    if (q.kp.size() >= 3) {                                 // B
                        if (listed == 0) len = (int)q.kp.size();
                        else if (len != (int)q.kp.size()) len = -1;
                        if ((size_t)listed < NP) {
                            float *dst = o + (size_t)listed * L;
                            if (off + L <= q.kp.size()) {
                                for (size_t j = 0; j < L; ++j) {
                                    const double v = q.kp[off + j];
                                    const float fv = (float)v;
                                    dst[j] = fv;
                                    my_inexact += (std::isfinite(v) && (double)fv != v);
                                }
                            }                                               // a short slice stays NaN (:203-205)
                        }
                        ++listed;
                    }
                }
                count_named[i] = named; count_listed[i] = listed; list_len[i] = len;
                status[i] = (named > max_persons || listed > max_persons) ? 3 : 1;
            }
        }
        inexact.fetch_add(my_inexact);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (n_inexact) *n_inexact = inexact.load();
    return P2S_OK;
}

namespace {

char *repr_double(char *p, double v);

// Re-emits a JSON document the way Python's `json.dumps(json.load(f))` writes it (separators ", " and ": ", floats as
// repr, "-0" -> "0", key order kept), with the elements of the top-level "people" array collected separately so that the
// caller can put the chosen ones in their place.  Gives up (`fallback`) on anything whose Python round trip is not a
// plain copy: strings with escapes or non-ASCII bytes, duplicate keys, a top level that is not an object.
class Reserializer {
  public:
    Reserializer(const char *b, const char *e) : p_(b), end_(e) {}
    bool fallback = false;
    bool has_people = false;
    size_t people_pos = 0;
    std::vector<std::string> people;

    bool run(std::string &out) {
        ws();
        if (p_ >= end_ || *p_ != '{') { fallback = true; return false; }
        if (!object(out, true)) return false;
        ws();
        return p_ == end_;
    }

  private:
    const char *p_, *end_;
    int depth_ = 0;
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool lit(const char *s, std::string &out) {
        const size_t n = std::strlen(s);
        if ((size_t)(end_ - p_) < n || std::memcmp(p_, s, n) != 0) return false;
        out.append(s, n);
        p_ += n;
        return true;
    }
    bool string(std::string &out, std::string *key) {
        if (p_ >= end_ || *p_ != '"') return false;
        const char *b = ++p_;
        while (p_ < end_ && *p_ != '"') {
            const unsigned char c = (unsigned char)*p_;
            if (c < 0x20 || c >= 0x7f || c == '\\') { fallback = true; return false; }
            ++p_;
        }
        if (p_ >= end_) return false;
        out.push_back('"'); out.append(b, p_); out.push_back('"');
        if (key) key->assign(b, p_);
        ++p_;
        return true;
    }
    bool number(std::string &out) {
        const char *s = p_;
        bool is_float = false;
        if (p_ < end_ && *p_ == '-') ++p_;
        if (p_ >= end_) return false;
        if (*p_ == '0') ++p_;
        else if (*p_ >= '1' && *p_ <= '9') { while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
        else return false;
        if (p_ < end_ && *p_ == '.') {
            is_float = true;
            ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            is_float = true;
            ++p_;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (!is_float) {
            if (p_ - s == 2 && s[0] == '-' && s[1] == '0') out.push_back('0');     // int("-0") == 0
            else out.append(s, p_);
            return true;
        }
        double v = 0.0;
        auto r = std::from_chars(s, p_, v);
        if (r.ec == std::errc::result_out_of_range) v = std::strtod(std::string(s, p_).c_str(), nullptr);
        else if (r.ec != std::errc()) return false;
        if (std::isinf(v)) { out.append(v < 0 ? "-Infinity" : "Infinity"); return true; }
        char tmp[40];
        char *e = repr_double(tmp, v);
        out.append(tmp, e);
        return true;
    }
    bool value(std::string &out) {
        ws();
        if (p_ >= end_) return false;
        const char c = *p_;
        if (c == '{') return object(out, false);
        if (c == '[') return array(out);
        if (c == '"') return string(out, nullptr);
        if (c == 't') return lit("true", out);
        if (c == 'f') return lit("false", out);
        if (c == 'n') return lit("null", out);
        if (c == 'N') return lit("NaN", out);
        if (c == 'I') return lit("Infinity", out);
        if (c == '-' && p_ + 1 < end_ && p_[1] == 'I') return lit("-Infinity", out);
        return number(out);
    }
    bool array(std::string &out) {
        if (++depth_ > 512) return false;
        ++p_;
        out.push_back('[');
        ws();
        if (p_ < end_ && *p_ == ']') { ++p_; --depth_; out.push_back(']'); return true; }
        for (;;) {
            if (!value(out)) return false;
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; out.append(", "); continue; }
            if (*p_ == ']') { ++p_; --depth_; out.push_back(']'); return true; }
            return false;
        }
    }
    bool people_array() {
        ++p_;
        ws();
        if (p_ < end_ && *p_ == ']') { ++p_; return true; }
        for (;;) {
            people.emplace_back();
            if (!value(people.back())) return false;
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; return true; }
            return false;
        }
    }
    bool object(std::string &out, bool top) {
        if (++depth_ > 512) return false;
        ++p_;
        out.push_back('{');
        ws();
        if (p_ < end_ && *p_ == '}') { ++p_; --depth_; out.push_back('}'); return true; }
        std::vector<std::string> keys;
        std::string key;
        for (;;) {
            ws();
            if (!string(out, &key)) return false;
            for (const std::string &k : keys) if (k == key) { fallback = true; return false; }   // dict: last value, first place
            keys.push_back(key);
            ws();
            if (p_ >= end_ || *p_ != ':') return false;
            ++p_;
            out.append(": ");
            ws();
            if (top && key == "people") {
                if (p_ >= end_ || *p_ != '[') { fallback = true; return false; }
                has_people = true;
                people_pos = out.size();
                if (!people_array()) return false;
            } else if (!value(out)) {
                return false;
            }
            ws();
            if (p_ >= end_) return false;
            if (*p_ == ',') { ++p_; out.append(", "); continue; }
            if (*p_ == '}') { ++p_; --depth_; out.push_back('}'); return true; }
            return false;
        }
    }
};

}  // namespace

// personAssociation.py:552-580 `rewrite_json_files` for all frames: per (frame, camera) the source document re-emitted
// with `people` replaced by the chosen person of every proposal of the frame ({} where the camera is off).  comb holds
// one row of n_cams indices per proposal (-1 = off), the proposals of frame f are rows prop_offset[f] .. prop_offset[f+1].
// A camera without a readable source, or an index past its people list, gets NO file (a stale one is removed), like the
// reference's `except: os.remove`.  status: 1 written, 0 no file, 2 left to the Python path (see Reserializer).
extern "C" int p2s_rewrite_people_files(const char *const *src, const char *const *dst, long long n_frames, int n_cams,
                                        const int32_t *prop_offset, const int32_t *comb, uint8_t *status, int n_threads) {
    if (!src || !dst || n_frames < 0 || n_cams < 1 || !prop_offset || !status) return P2S_EINVAL;
    const long long n_files = n_frames * n_cams;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((long long)n_threads > n_files) n_threads = (int)(n_files > 0 ? n_files : 1);
    std::atomic<long long> next{0};
    auto work = [&]() {
        std::string buf, out, fin;
        for (;;) {
            const long long i0 = next.fetch_add(16);
            if (i0 >= n_files) break;
            const long long i1 = i0 + 16 < n_files ? i0 + 16 : n_files;
            for (long long i = i0; i < i1; ++i) {
                const long long f = i / n_cams;
                const int c = (int)(i % n_cams);
                status[i] = 0;
                if (!dst[i] || !dst[i][0]) continue;
                bool ok = src[i] && src[i][0] && read_file(src[i], buf);
                if (ok) {
                    out.clear();
                    Reserializer rs(buf.data(), buf.data() + buf.size());
                    ok = rs.run(out);
                    if (rs.fallback || (ok && !rs.has_people)) { status[i] = 2; continue; }
                    if (ok) {
                        fin.assign(out, 0, rs.people_pos);
                        fin.push_back('[');
                        for (int32_t r = prop_offset[f]; r < prop_offset[f + 1] && ok; ++r) {
                            const int32_t idx = comb[(size_t)r * n_cams + c];
                            if (r > prop_offset[f]) fin.append(", ");
                            if (idx < 0) fin.append("{}");
                            else if ((size_t)idx < rs.people.size()) fin.append(rs.people[(size_t)idx]);
                            else ok = false;                                   // IndexError in the reference
                        }
                        fin.push_back(']');
                        fin.append(out, rs.people_pos, std::string::npos);
                    }
                }
                if (ok) {
                    FILE *g = std::fopen(dst[i], "wb");
                    ok = g != nullptr;
                    if (g) {
                        ok = std::fwrite(fin.data(), 1, fin.size(), g) == fin.size();
                        ok = (std::fclose(g) == 0) && ok;
                    }
                }
                if (ok) status[i] = 1;
                else std::remove(dst[i]);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return P2S_OK;
}

// ---- TRC body writer (Pose2Sim/triangulation.py:206-213: DataFrame.to_csv(sep='\t', header=None)) -----------
// One row per frame: frame number, time, 3 K coordinates, tab separated, '\n' terminated; NaN is an empty
// field (pandas na_rep='').  Numbers are written like Python's repr(float) — the shortest string that
// round-trips, fixed notation for 1e-4 <= |v| < 1e16 (always with a fractional part, "2.0"), scientific
// with a two-digit exponent otherwise ("1e-05") — which is what pandas emits for float64 columns.
namespace {

char *repr_double(char *p, double v) {
    if (std::isnan(v)) return p;                                   // empty field
    if (std::isinf(v)) { const char *s = v > 0 ? "inf" : "-inf"; while (*s) *p++ = *s++; return p; }
    if (v == 0.0) { if (std::signbit(v)) *p++ = '-'; *p++ = '0'; *p++ = '.'; *p++ = '0'; return p; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);   // shortest: d[.ddd]e[+-]XX
    char *b = buf, *e = r.ptr;
    if (*b == '-') { *p++ = '-'; ++b; }
    char digits[24];
    int nd = 0;
    char *q = b;
    for (; q < e && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
    int exp10 = 0;
    {
        ++q;                                                        // past 'e'
        const bool neg = (*q == '-');
        ++q;
        for (; q < e; ++q) exp10 = exp10 * 10 + (*q - '0');
        if (neg) exp10 = -exp10;
    }
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {                                            // 0.000ddd
            *p++ = '0'; *p++ = '.';
            for (int i = 0; i < -exp10 - 1; ++i) *p++ = '0';
            for (int i = 0; i < nd; ++i) *p++ = digits[i];
        } else {
            int i = 0;
            for (; i <= exp10; ++i) *p++ = (i < nd) ? digits[i] : '0';
            *p++ = '.';
            if (i >= nd) *p++ = '0';
            for (; i < nd; ++i) *p++ = digits[i];
        }
    } else {                                                        // d[.ddd]e[+-]XX
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; for (int i = 1; i < nd; ++i) *p++ = digits[i]; }
        *p++ = 'e';
        *p++ = exp10 < 0 ? '-' : '+';
        const int a = exp10 < 0 ? -exp10 : exp10;
        if (a >= 100) { *p++ = (char)('0' + a / 100); *p++ = (char)('0' + (a / 10) % 10); *p++ = (char)('0' + a % 10); }
        else { *p++ = (char)('0' + a / 10); *p++ = (char)('0' + a % 10); }
    }
    return p;
}

}  // namespace

// The same rows formatted into a caller-provided buffer (frame-block sharded writers: every rank formats its own rows,
// the ranks exchange byte counts and write their ranges of the one file in parallel).  cap >= n_rows * (22 + 25 *
// (n_cols + 1)) always suffices.
extern "C" int p2s_format_trc_rows(const long long *frames, const double *time_s, const double *values, long long n_rows,
                                   int n_cols, char *buf, size_t cap, size_t *len) {
    if (!len || n_rows < 0 || n_cols < 0 || (n_rows > 0 && (!buf || !frames || !time_s || (n_cols > 0 && !values)))) return P2S_EINVAL;
    char *p = buf;
    const size_t per_row = 22 + 25 * ((size_t)n_cols + 1);
    for (long long r = 0; r < n_rows; ++r) {
        if ((size_t)(p - buf) + per_row > cap) return P2S_EINVAL;
        auto fr = std::to_chars(p, p + 24, frames[r]);
        p = fr.ptr;
        *p++ = '\t';
        p = repr_double(p, time_s[r]);
        const double *row = values + (size_t)r * (size_t)n_cols;
        for (int c = 0; c < n_cols; ++c) { *p++ = '\t'; p = repr_double(p, row[c]); }
        *p++ = '\n';
    }
    *len = (size_t)(p - buf);
    return P2S_OK;
}

extern "C" int p2s_write_trc_rows(const char *path, const long long *frames, const double *time_s,
                                  const double *values, long long n_rows, int n_cols) {
    if (!path || n_rows < 0 || n_cols < 0 || (n_rows > 0 && (!frames || !time_s || (n_cols > 0 && !values)))) return P2S_EINVAL;
    FILE *f = std::fopen(path, "ab");
    if (!f) return P2S_EINVAL;
    // rows are formatted in parallel (shortest round-trip conversion is ~100 ns per number: 5000 frames x 80 columns
    // would be 50 ms on one thread), one contiguous block of rows per thread, then written in order
    int n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((long long)n_threads > (n_rows + 255) / 256) n_threads = (int)((n_rows + 255) / 256);
    if (n_threads < 1) n_threads = 1;
    const size_t per_row = 22 + 25 * ((size_t)n_cols + 1);
    std::vector<std::vector<char>> bufs((size_t)n_threads);
    std::vector<size_t> lens((size_t)n_threads, 0);
    std::vector<int> rcs((size_t)n_threads, P2S_OK);
    auto work = [&](int t) {
        const long long r0 = n_rows * t / n_threads, r1 = n_rows * (t + 1) / n_threads;
        bufs[(size_t)t].resize((size_t)(r1 - r0) * per_row + 1);
        rcs[(size_t)t] = p2s_format_trc_rows(frames + r0, time_s + r0, values ? values + (size_t)r0 * (size_t)n_cols : nullptr, r1 - r0,
                                              n_cols, bufs[(size_t)t].data(), bufs[(size_t)t].size(), &lens[(size_t)t]);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &t : pool) t.join();
    bool ok = true;
    for (int t = 0; t < n_threads && ok; ++t) {
        ok = rcs[(size_t)t] == P2S_OK;
        if (ok && lens[(size_t)t]) ok = std::fwrite(bufs[(size_t)t].data(), 1, lens[(size_t)t], f) == lens[(size_t)t];
    }
    ok = !std::ferror(f) && ok;
    std::fclose(f);
    return ok ? P2S_OK : P2S_EINVAL;
}
