// obj_format.cu - Wavefront OBJ text of an extracted mesh, byte-identical to the reference's writer.
//
// Replaces the per-line Python loops of write_obj (model/render/obj.py:128-177).  The reference formats every coordinate
// with '{}'.format(np.float32), which numpy hands to float.__format__: the text is Python's repr of the value WIDENED TO
// DOUBLE (shortest decimal string that round-trips the double, e.g. np.float32(0.1) -> "0.10000000149011612"), and the
// texcoord v is flipped in float32 first (`1.0 - v[1]`, :148).  Host-only code (the text has to reach a file): a
// shortest-round-trip double formatter (Ryu-style: 128-bit fixed-point powers of 5, tables generated at first use with
// exact integer arithmetic) + Python's repr layout rules, run over fixed-size line chunks by all host threads, then
// compacted in place.  No CUDA in this file; it lives in libb2a.so so the same C-ABI serves it.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------------------------
// Tables: POW5[i] = the top 125 bits of 5^i (floor), INV5[i] = floor(2^(len(5^i) - 1 + 125) / 5^i) + 1.
// ---------------------------------------------------------------------------------------------------------------
const int kPow5Bits = 125;
const int kPow5Count = 326;
const int kInv5Count = 342;
uint64_t g_pow5[kPow5Count][2];
uint64_t g_inv5[kInv5Count][2];
std::once_flag g_tables_once;

const int kLimbs = 32;                                   // 1024 bits: 2^(795 + 125) fits
struct Big {
    uint32_t w[kLimbs];
};

int big_bitlen(const Big& a)
{
    for (int i = kLimbs - 1; i >= 0; --i)
        if (a.w[i]) return 32 * i + 32 - __builtin_clz(a.w[i]);
    return 0;
}

void big_mul_small(Big& a, uint32_t m)
{
    uint64_t carry = 0;
    for (int i = 0; i < kLimbs; ++i) {
        uint64_t t = (uint64_t)a.w[i] * m + carry;
        a.w[i] = (uint32_t)t;
        carry = t >> 32;
    }
}

void big_div_small(Big& a, uint32_t d)                   // floor division in place
{
    uint64_t rem = 0;
    for (int i = kLimbs - 1; i >= 0; --i) {
        uint64_t t = (rem << 32) | a.w[i];
        a.w[i] = (uint32_t)(t / d);
        rem = t % d;
    }
}

int big_bit(const Big& a, int b) { return (b < 0 || b >= 32 * kLimbs) ? 0 : (a.w[b >> 5] >> (b & 31)) & 1; }

// bits [lo, lo + 128) of a (lo may be negative: zeros enter from below)
void big_extract128(const Big& a, int lo, uint64_t out[2])
{
    out[0] = out[1] = 0;
    for (int b = 0; b < 128; ++b)
        if (big_bit(a, lo + b)) out[b >> 6] |= 1ull << (b & 63);
}

void build_tables()
{
    Big p;
    memset(&p, 0, sizeof(p));
    p.w[0] = 1;
    for (int i = 0; i < kInv5Count; ++i) {
        int len = big_bitlen(p);
        if (i < kPow5Count) big_extract128(p, len - kPow5Bits, g_pow5[i]);
        // floor(floor(x / 5) / 5) = floor(x / 25): i short divisions of 2^j give floor(2^j / 5^i) exactly
        Big q;
        memset(&q, 0, sizeof(q));
        int j = len - 1 + kPow5Bits;
        q.w[j >> 5] = 1u << (j & 31);
        for (int k = 0; k < i; ++k) big_div_small(q, 5);
        big_extract128(q, 0, g_inv5[i]);
        if (++g_inv5[i][0] == 0) ++g_inv5[i][1];
        big_mul_small(p, 5);
    }
}

inline int pow5bits(int e) { return (int)(((uint32_t)e * 1217359u) >> 19) + 1; }     // bit length of 5^e, 0 <= e <= 3528
inline int log10_pow2(int e) { return (int)(((uint32_t)e * 78913u) >> 18); }          // floor(log10(2^e)), 0 <= e <= 1650
inline int log10_pow5(int e) { return (int)(((uint32_t)e * 732923u) >> 20); }         // floor(log10(5^e)), 0 <= e <= 2620

inline uint64_t mul_shift(uint64_t m, const uint64_t* mul, int j)                     // (m * mul) >> j, 64 <= j < 128
{
    u128 lo = (u128)m * mul[0];
    u128 hi = (u128)m * mul[1];
    return (uint64_t)(((lo >> 64) + hi) >> (j - 64));
}

inline bool multiple_of_pow5(uint64_t v, int p)
{
    int n = 0;
    while (v % 5 == 0) {
        v /= 5;
        if (++n >= p) return true;
    }
    return n >= p;
}

inline bool multiple_of_pow2(uint64_t v, int p) { return (v & ((1ull << p) - 1)) == 0; }

// Shortest decimal (digits, exponent) that rounds to the finite non-zero double with these IEEE fields: round-to-even
// interval bounds, closest candidate among the shortest (the string David Gay's mode-0 dtoa, i.e. Python's repr, produces).
void shortest_decimal(uint64_t ieee_mantissa, int ieee_exponent, uint64_t* digits, int* exponent)
{
    int e2;
    uint64_t m2;
    if (ieee_exponent == 0) {
        e2 = 1 - 1023 - 52 - 2;
        m2 = ieee_mantissa;
    } else {
        e2 = ieee_exponent - 1023 - 52 - 2;
        m2 = (1ull << 52) | ieee_mantissa;
    }
    const bool accept = (m2 & 1) == 0;
    const uint64_t mv = 4 * m2;
    const uint32_t mm_shift = (ieee_mantissa != 0 || ieee_exponent <= 1) ? 1 : 0;
    uint64_t vr, vp, vm;
    int e10;
    bool vm_tz = false, vr_tz = false;
    if (e2 >= 0) {
        const int q = log10_pow2(e2) - (e2 > 3);
        e10 = q;
        const int k = kPow5Bits + pow5bits(q) - 1;
        const int i = -e2 + q + k;
        vr = mul_shift(4 * m2, g_inv5[q], i);
        vp = mul_shift(4 * m2 + 2, g_inv5[q], i);
        vm = mul_shift(4 * m2 - 1 - mm_shift, g_inv5[q], i);
        if (q <= 21) {
            if (mv % 5 == 0)
                vr_tz = multiple_of_pow5(mv, q);
            else if (accept)
                vm_tz = multiple_of_pow5(mv - 1 - mm_shift, q);
            else
                vp -= multiple_of_pow5(mv + 2, q);
        }
    } else {
        const int q = log10_pow5(-e2) - (-e2 > 1);
        e10 = q + e2;
        const int i = -e2 - q;
        const int k = pow5bits(i) - kPow5Bits;
        const int j = q - k;
        vr = mul_shift(4 * m2, g_pow5[i], j);
        vp = mul_shift(4 * m2 + 2, g_pow5[i], j);
        vm = mul_shift(4 * m2 - 1 - mm_shift, g_pow5[i], j);
        if (q <= 1) {
            vr_tz = true;
            if (accept)
                vm_tz = mm_shift == 1;
            else
                --vp;
        } else if (q < 63) {
            vr_tz = multiple_of_pow2(mv, q);
        }
    }
    int removed = 0;
    uint32_t last = 0;
    uint64_t out;
    if (vm_tz || vr_tz) {
        while (vp / 10 > vm / 10) {
            vm_tz &= vm % 10 == 0;
            vr_tz &= last == 0;
            last = (uint32_t)(vr % 10);
            vr /= 10; vp /= 10; vm /= 10;
            ++removed;
        }
        if (vm_tz) {
            while (vm % 10 == 0) {
                vr_tz &= last == 0;
                last = (uint32_t)(vr % 10);
                vr /= 10; vp /= 10; vm /= 10;
                ++removed;
            }
        }
        if (vr_tz && last == 5 && vr % 2 == 0) last = 4;          // exactly half: round to even
        out = vr + (((vr == vm) && (!accept || !vm_tz)) || last >= 5);
    } else {
        bool up = false;
        while (vp / 10 > vm / 10) {
            up = vr % 10 >= 5;
            vr /= 10; vp /= 10; vm /= 10;
            ++removed;
        }
        out = vr + (vr == vm || up);
    }
    *digits = out;
    *exponent = e10 + removed;
}

// Python's repr(float) for the double `x`: positional for -4 < decimal point position <= 16, else d.ddde+XX.
inline char* write_repr(char* p, double x)
{
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const uint64_t mant = bits & ((1ull << 52) - 1);
    const int expo = (int)((bits >> 52) & 0x7ff);
    if (expo == 0x7ff) {
        if (mant) { memcpy(p, "nan", 3); return p + 3; }
        if (bits >> 63) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    if (bits >> 63) *p++ = '-';
    if (expo == 0 && mant == 0) { memcpy(p, "0.0", 3); return p + 3; }
    uint64_t d;
    int e;
    shortest_decimal(mant, expo, &d, &e);
    char buf[20];
    int n = 0;
    while (d) { buf[n++] = (char)('0' + d % 10); d /= 10; }      // reversed
    const int decpt = e + n;
    if (decpt > 16 || decpt <= -4) {
        *p++ = buf[n - 1];
        if (n > 1) {
            *p++ = '.';
            for (int i = n - 2; i >= 0; --i) *p++ = buf[i];
        }
        *p++ = 'e';
        int ex = decpt - 1;
        *p++ = ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        if (ex >= 100) { *p++ = (char)('0' + ex / 100); ex %= 100; }
        *p++ = (char)('0' + ex / 10);
        *p++ = (char)('0' + ex % 10);
    } else if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        for (int i = n - 1; i >= 0; --i) *p++ = buf[i];
    } else if (decpt >= n) {
        for (int i = n - 1; i >= 0; --i) *p++ = buf[i];
        for (int i = n; i < decpt; ++i) *p++ = '0';
        *p++ = '.'; *p++ = '0';
    } else {
        for (int i = n - 1; i >= n - decpt; --i) *p++ = buf[i];
        *p++ = '.';
        for (int i = n - decpt - 1; i >= 0; --i) *p++ = buf[i];
    }
    return p;
}

inline char* write_int(char* p, int64_t v)
{
    uint64_t u = (uint64_t)v;
    if (v < 0) { *p++ = '-'; u = 0 - u; }
    char buf[20];
    int n = 0;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (n) *p++ = buf[--n];
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// Sections of the file, each a run of fixed-upper-bound lines
// ---------------------------------------------------------------------------------------------------------------
enum { SEC_V = 0, SEC_VT = 1, SEC_VN = 2, SEC_F = 3 };
const size_t kLineBound[4] = {80, 56, 80, 192};       // 'v ' + 3 x (24-char repr + ' ') + '\n' etc.; faces: 9 x 20-digit ints
const int64_t kChunkLines = 2048;

struct Mesh {
    const float* v_pos;
    const float* v_tex;
    const float* v_nrm;
    const int64_t* t_pos;
    const int64_t* t_tex;
    const int64_t* t_nrm;
};

size_t format_lines(const Mesh& m, int sec, int64_t lo, int64_t hi, char* out)
{
    char* p = out;
    for (int64_t i = lo; i < hi; ++i) {
        if (sec == SEC_V) {                                               // 'v {} {} {} \n' (obj.py:143)
            *p++ = 'v';
            for (int c = 0; c < 3; ++c) { *p++ = ' '; p = write_repr(p, (double)m.v_pos[3 * i + c]); }
            *p++ = ' ';
        } else if (sec == SEC_VT) {                                       // 'vt {} {} \n', v flipped in float32 (:148)
            *p++ = 'v'; *p++ = 't'; *p++ = ' ';
            p = write_repr(p, (double)m.v_tex[2 * i]);
            *p++ = ' ';
            volatile float flipped = 1.0f - m.v_tex[2 * i + 1];
            p = write_repr(p, (double)flipped);
            *p++ = ' ';
        } else if (sec == SEC_VN) {                                       // 'vn {} {} {}\n' (:153)
            *p++ = 'v'; *p++ = 'n';
            for (int c = 0; c < 3; ++c) { *p++ = ' '; p = write_repr(p, (double)m.v_nrm[3 * i + c]); }
        } else {                                                          // 'f ' + 3 x ' %s/%s/%s' (:163-166)
            *p++ = 'f'; *p++ = ' ';
            for (int c = 0; c < 3; ++c) {
                *p++ = ' ';
                p = write_int(p, m.t_pos[3 * i + c] + 1);
                *p++ = '/';
                if (m.t_tex) p = write_int(p, m.t_tex[3 * i + c] + 1);
                *p++ = '/';
                if (m.t_nrm) p = write_int(p, m.t_nrm[3 * i + c] + 1);
            }
        }
        *p++ = '\n';
    }
    return (size_t)(p - out);
}

struct Chunk {
    int sec;
    int64_t lo, hi;
    size_t offset;      // worst-case offset in the output buffer
    size_t bytes;       // actual bytes written there
};

}  // namespace

// worst-case size of the text for these counts (what the caller allocates)
B2A_API int b2a_obj_text_bound(int64_t n_pos, int64_t n_tex, int64_t n_nrm, int64_t n_faces, int64_t name_bytes, size_t* bytes)
{
    if (n_pos < 0 || n_tex < 0 || n_nrm < 0 || n_faces < 0 || name_bytes < 0 || !bytes) {
        b2a_set_error("b2a_obj_text_bound: bad arguments");
        return 1;
    }
    *bytes = 128 + (size_t)name_bytes + (size_t)n_pos * kLineBound[SEC_V] + (size_t)n_tex * kLineBound[SEC_VT] +
             (size_t)n_nrm * kLineBound[SEC_VN] + (size_t)n_faces * kLineBound[SEC_F];
    return 0;
}

// HOST pointers.  v_tex / v_nrm may be NULL (their index columns are then left empty, obj.py:166); n_tex = 0 with a
// non-NULL v_tex writes no 'vt' lines but keeps the texcoord indices (the reference's save_material=False case, :145).
B2A_API int b2a_obj_format(const float* v_pos, int64_t n_pos, const float* v_tex, int64_t n_tex, const float* v_nrm, int64_t n_nrm,
                           const int64_t* t_pos_idx, const int64_t* t_tex_idx, const int64_t* t_nrm_idx, int64_t n_faces,
                           const char* mtl_name, int64_t name_bytes, char* out, size_t capacity, size_t* written, int threads)
{
    size_t bound = 0;
    if (b2a_obj_text_bound(n_pos, n_tex, n_nrm, n_faces, name_bytes, &bound)) return 1;
    if (!out || !written || capacity < bound) {
        b2a_set_error("b2a_obj_format: output buffer of %zu bytes is smaller than the bound %zu", capacity, bound);
        return 1;
    }
    if ((n_pos && !v_pos) || (n_tex && !v_tex) || (n_nrm && !v_nrm) || (n_faces && !t_pos_idx) || (v_tex && n_faces && !t_tex_idx) ||
        (v_nrm && n_faces && !t_nrm_idx) || (name_bytes && !mtl_name)) {
        b2a_set_error("b2a_obj_format: NULL input with a non-zero count");
        return 1;
    }
    std::call_once(g_tables_once, build_tables);
    Mesh m = {v_pos, v_tex, v_nrm, t_pos_idx, v_tex ? t_tex_idx : nullptr, v_nrm ? t_nrm_idx : nullptr};

    char* p = out;                                                        // obj.py:132-133
    memcpy(p, "mtllib ", 7); p += 7;
    if (name_bytes) memcpy(p, mtl_name, (size_t)name_bytes);
    p += name_bytes;
    memcpy(p, ".mtl\ng default\n", 15); p += 15;
    const size_t head = (size_t)(p - out);

    static const char kFaceHead[] = "s 1 \ng pMesh1\nusemtl defaultMat\n";   // :157-159
    const size_t face_head = sizeof(kFaceHead) - 1;
    std::vector<Chunk> chunks;
    size_t off = head;
    const int64_t counts[4] = {n_pos, n_tex, v_nrm ? n_nrm : 0, n_faces};
    size_t face_head_at = 0;
    for (int sec = 0; sec < 4; ++sec) {
        if (sec == SEC_F) { face_head_at = chunks.size(); off += face_head; }
        for (int64_t lo = 0; lo < counts[sec]; lo += kChunkLines) {
            int64_t hi = lo + kChunkLines < counts[sec] ? lo + kChunkLines : counts[sec];
            chunks.push_back({sec, lo, hi, off, 0});
            off += (size_t)(hi - lo) * kLineBound[sec];
        }
    }
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((size_t)nt > chunks.size()) nt = chunks.size() ? (int)chunks.size() : 1;
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= chunks.size()) return;
            Chunk& ch = chunks[c];
            ch.bytes = format_lines(m, ch.sec, ch.lo, ch.hi, out + ch.offset);
        }
    };
    if (nt == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        try {
            for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        } catch (...) {
            // thread creation refused (resource limits): the threads that did start and this one finish the chunks
        }
        work();
        for (auto& th : pool) th.join();
    }
    // compaction: chunks only ever move towards the front
    p = out + head;
    for (size_t c = 0; c <= chunks.size(); ++c) {
        if (c == face_head_at) { memcpy(p, kFaceHead, face_head); p += face_head; }
        if (c == chunks.size()) break;
        memmove(p, out + chunks[c].offset, chunks[c].bytes);
        p += chunks[c].bytes;
    }
    *written = (size_t)(p - out);
    return 0;
}
