// Front-end results as one byte stream (include/sfe.h, "SFER"): host-only pack / unpack of the cap-strided arrays the
// batched entry points fill.  The reference has no such format (src/pipeline.cpp:231-241 is `#if 0`); SURVEY §8f row 4.
#include <algorithm>
#include <cstring>

#include "sfe_common.cuh"

namespace {

constexpr uint32_t kMagic = 0x52454653u;  // "SFER"
constexpr uint32_t kVersion = 1;
constexpr size_t kHeader = 48;

uint64_t fnv1a64(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

size_t frame_bytes(int nl, int nr, int flags) {
    size_t b = 8 + (size_t)nl * (28 + 32);
    if (flags & SFE_RES_STEREO) b += (size_t)nr * (28 + 32) + (size_t)nl * 8;
    if (flags & SFE_RES_TRACK) b += (size_t)nl * 8;
    return b;
}

template <typename T>
void put(uint8_t *&p, const T &v) {
    memcpy(p, &v, sizeof(T));
    p += sizeof(T);
}
template <typename T>
T get(const uint8_t *&p) {
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
}
void put_rows(uint8_t *&p, const void *src, size_t bytes) {
    memcpy(p, src, bytes);
    p += bytes;
}
void get_rows(const uint8_t *&p, void *dst, size_t bytes) {
    if (dst) memcpy(dst, p, bytes);
    p += bytes;
}

struct Header {
    uint32_t frames, flags, w, h;
    uint64_t payload, sum;
};

int parse_header(const void *buf, size_t bytes, Header &H) {
    SFE_REQUIRE(buf && bytes >= kHeader, SFE_ERR_BAD_ARG, "buffer shorter than the header");
    const uint8_t *p = (const uint8_t *)buf;
    SFE_REQUIRE(get<uint32_t>(p) == kMagic, SFE_ERR_BAD_ARG, "not an SFER stream");
    SFE_REQUIRE(get<uint32_t>(p) == kVersion, SFE_ERR_UNSUPPORTED, "unknown SFER version");
    H.frames = get<uint32_t>(p);
    H.flags = get<uint32_t>(p);
    H.w = get<uint32_t>(p);
    H.h = get<uint32_t>(p);
    H.payload = get<uint64_t>(p);
    H.sum = get<uint64_t>(p);
    SFE_REQUIRE(H.frames <= (1u << 24) && (H.flags & ~3u) == 0, SFE_ERR_BAD_ARG, "corrupt header");
    SFE_REQUIRE(H.payload == bytes - kHeader, SFE_ERR_BAD_ARG, "stream length does not match the header");
    SFE_REQUIRE(fnv1a64((const uint8_t *)buf + kHeader, (size_t)H.payload) == H.sum, SFE_ERR_BAD_ARG, "checksum mismatch");
    return SFE_OK;
}

}  // namespace

extern "C" {

int sfe_results_size(int frames, const int32_t *n_l, const int32_t *n_r, int flags, size_t *bytes) {
    SFE_REQUIRE(frames >= 0 && bytes && (frames == 0 || n_l) && (flags & ~3) == 0, SFE_ERR_BAD_ARG, "bad argument");
    SFE_REQUIRE(!(flags & SFE_RES_STEREO) || frames == 0 || n_r, SFE_ERR_BAD_ARG, "stereo stream without n_r");
    size_t b = kHeader;
    for (int f = 0; f < frames; f++) {
        const int nr = (flags & SFE_RES_STEREO) ? n_r[f] : 0;
        SFE_REQUIRE(n_l[f] >= 0 && nr >= 0, SFE_ERR_BAD_ARG, "negative keypoint count");
        b += frame_bytes(n_l[f], nr, flags);
    }
    *bytes = b;
    return SFE_OK;
}

int sfe_results_pack(void *buf, size_t buf_bytes, int frames, int cap, int w, int h, int flags, const sfe_keypoint *kps_l,
                     const uint8_t *desc_l, const int32_t *n_l, const sfe_keypoint *kps_r, const uint8_t *desc_r,
                     const int32_t *n_r, const int32_t *stereo_idx, const int32_t *stereo_dist, const int32_t *track_idx,
                     const int32_t *track_dist, size_t *written) {
    size_t need = 0;
    if (int rc = sfe_results_size(frames, n_l, n_r, flags, &need)) return rc;
    SFE_REQUIRE(buf && buf_bytes >= need, SFE_ERR_CAPACITY, "buffer smaller than sfe_results_size()");
    SFE_REQUIRE(cap >= 0 && w >= 0 && h >= 0 && (frames == 0 || (kps_l && desc_l)), SFE_ERR_BAD_ARG, "bad argument");
    SFE_REQUIRE(!(flags & SFE_RES_STEREO) || frames == 0 || (kps_r && desc_r && stereo_idx && stereo_dist), SFE_ERR_BAD_ARG,
                "stereo stream without its arrays");
    SFE_REQUIRE(!(flags & SFE_RES_TRACK) || frames == 0 || (track_idx && track_dist), SFE_ERR_BAD_ARG, "track stream without its arrays");
    uint8_t *p = (uint8_t *)buf + kHeader;
    for (int f = 0; f < frames; f++) {
        const int nl = n_l[f], nr = (flags & SFE_RES_STEREO) ? n_r[f] : 0;
        SFE_REQUIRE(nl <= cap && nr <= cap, SFE_ERR_BAD_ARG, "keypoint count exceeds cap");
        const size_t o = (size_t)f * cap;
        put<uint32_t>(p, (uint32_t)nl);
        put<uint32_t>(p, (uint32_t)nr);
        put_rows(p, kps_l + o, (size_t)nl * 28);
        put_rows(p, desc_l + o * 32, (size_t)nl * 32);
        if (flags & SFE_RES_STEREO) {
            put_rows(p, kps_r + o, (size_t)nr * 28);
            put_rows(p, desc_r + o * 32, (size_t)nr * 32);
            put_rows(p, stereo_idx + o, (size_t)nl * 4);
            put_rows(p, stereo_dist + o, (size_t)nl * 4);
        }
        if (flags & SFE_RES_TRACK) {
            put_rows(p, track_idx + o, (size_t)nl * 4);
            put_rows(p, track_dist + o, (size_t)nl * 4);
        }
    }
    const uint64_t payload = (uint64_t)(p - ((uint8_t *)buf + kHeader));
    uint8_t *q = (uint8_t *)buf;
    put<uint32_t>(q, kMagic);
    put<uint32_t>(q, kVersion);
    put<uint32_t>(q, (uint32_t)frames);
    put<uint32_t>(q, (uint32_t)flags);
    put<uint32_t>(q, (uint32_t)w);
    put<uint32_t>(q, (uint32_t)h);
    put<uint64_t>(q, payload);
    put<uint64_t>(q, fnv1a64((const uint8_t *)buf + kHeader, (size_t)payload));
    put<uint64_t>(q, 0);
    if (written) *written = kHeader + (size_t)payload;
    return SFE_OK;
}

int sfe_results_info(const void *buf, size_t bytes, int *frames, int *flags, int *w, int *h, int *max_n) {
    Header H;
    if (int rc = parse_header(buf, bytes, H)) return rc;
    const uint8_t *p = (const uint8_t *)buf + kHeader, *end = (const uint8_t *)buf + bytes;
    int mx = 0;
    for (uint32_t f = 0; f < H.frames; f++) {
        SFE_REQUIRE((size_t)(end - p) >= 8, SFE_ERR_BAD_ARG, "truncated frame record");
        const uint32_t nl = get<uint32_t>(p), nr = get<uint32_t>(p);
        SFE_REQUIRE(nl < (1u << 24) && nr < (1u << 24), SFE_ERR_BAD_ARG, "corrupt frame record");
        const size_t b = frame_bytes((int)nl, (int)nr, (int)H.flags) - 8;
        SFE_REQUIRE((size_t)(end - p) >= b, SFE_ERR_BAD_ARG, "truncated frame record");
        p += b;
        mx = std::max(mx, (int)std::max(nl, nr));
    }
    SFE_REQUIRE(p == end, SFE_ERR_BAD_ARG, "trailing bytes after the last frame");
    if (frames) *frames = (int)H.frames;
    if (flags) *flags = (int)H.flags;
    if (w) *w = (int)H.w;
    if (h) *h = (int)H.h;
    if (max_n) *max_n = mx;
    return SFE_OK;
}

int sfe_results_unpack(const void *buf, size_t bytes, int cap, sfe_keypoint *kps_l, uint8_t *desc_l, int32_t *n_l,
                       sfe_keypoint *kps_r, uint8_t *desc_r, int32_t *n_r, int32_t *stereo_idx, int32_t *stereo_dist,
                       int32_t *track_idx, int32_t *track_dist) {
    int frames = 0, flags = 0, mx = 0;
    if (int rc = sfe_results_info(buf, bytes, &frames, &flags, nullptr, nullptr, &mx)) return rc;
    SFE_REQUIRE(cap >= mx, SFE_ERR_CAPACITY, "cap smaller than the largest frame");
    const uint8_t *p = (const uint8_t *)buf + kHeader;
    for (int f = 0; f < frames; f++) {
        const int nl = (int)get<uint32_t>(p), nr = (int)get<uint32_t>(p);
        const size_t o = (size_t)f * cap;
        if (n_l) n_l[f] = nl;
        if (n_r) n_r[f] = nr;
        get_rows(p, kps_l ? kps_l + o : nullptr, (size_t)nl * 28);
        get_rows(p, desc_l ? desc_l + o * 32 : nullptr, (size_t)nl * 32);
        if (flags & SFE_RES_STEREO) {
            get_rows(p, kps_r ? kps_r + o : nullptr, (size_t)nr * 28);
            get_rows(p, desc_r ? desc_r + o * 32 : nullptr, (size_t)nr * 32);
            get_rows(p, stereo_idx ? stereo_idx + o : nullptr, (size_t)nl * 4);
            get_rows(p, stereo_dist ? stereo_dist + o : nullptr, (size_t)nl * 4);
        }
        if (flags & SFE_RES_TRACK) {
            get_rows(p, track_idx ? track_idx + o : nullptr, (size_t)nl * 4);
            get_rows(p, track_dist ? track_dist + o : nullptr, (size_t)nl * 4);
        }
    }
    return SFE_OK;
}

}  // extern "C"
