// geomc_b200/batch.hpp -- C++ host side of the drop-in: the reference's geomc/linalg factor /
// solve / inverse functions, one call per BATCH, same names, argument meaning and error behaviour
// (bool / degenerate-count results per system, inputs destroyed where the reference destroys
// them, outputs untouched where the reference leaves them untouched).
//
// Builds with a plain C++17 compiler; links libgeomc_b200.so (the C ABI of ../geomc_b200.h).
// There is no CPU fallback: every call ends in a CUDA kernel launch or returns an error.
//
//   namespace geom::batch
//     BatchSoA<T>                      the SoA-interleaved batch container (replaces T[B][E] of objects)
//     decomp_plu / decomp_lup          LUDecomp.h:188 / :102
//     backsolve_lu / backsolve_plu     LUDecomp.h:332 / :271
//     linear_solve (pointer form)      LUDecomp.h:388
//     linear_solve (Vec form)          LUDecomp.h:417
//     cholesky                         Cholesky.h:37 / :83
//     cholesky_solve                   (new; the reference has no Cholesky solve)
//     inv                              Matrix.h:717 -> MatrixInv.h:49-337
//     orthogonal / nullspace           Orthogonal.h:73 / :136
//     contains / trace_simplex / intersect      shape/Simplex.h:450 / :647 / :251
//     mul / mul_acc                    Matrix.h:386 / :433
//     transformation / translation / scale / compose / apply...   AffineTransform.h:862 / :814 / :838 / :142 / :224-293
//
// Three forms of every function:
//   (1) device form  -- DeviceBatch / BatchSoA arguments already resident in HBM, asynchronous;
//   (2) host form    -- raw host arrays `T a[B][n*n]` exactly as a loop over the reference's
//                       pointer functions would see them; copies in, solves, copies out;
//   (3) object form  -- arrays of the reference's own objects (geom::SimpleMatrix<T,N,N,Lyt>,
//                       geom::Vec<T,N>), duck-typed so this header does not include geomc.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../geomc_b200.h"

namespace geom {
namespace batch {

using index_t = std::ptrdiff_t;          // geomc_defs.h:85

enum class Layout : int { AoS = GMB_LAYOUT_AOS, SoA = GMB_LAYOUT_SOA };

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* what) : std::runtime_error(std::string(what) + ": " + gmb_error_string(c)), code(c) {}
};
inline void check(int rc, const char* what) {
    if (rc != GMB_OK) throw Error(rc, what);
}

// ---------------------------------------------------------------- type -> C ABI entry points
template <typename T> struct Abi;
#define GMB_ABI(T, S)                                                                                       \
    template <> struct Abi<T> {                                                                             \
        static constexpr auto aos_to_soa = gmb_aos_to_soa_##S;                                              \
        static constexpr auto soa_to_aos = gmb_soa_to_aos_##S;                                              \
        static constexpr auto plu_decomp = gmb_plu_decomp_##S;                                              \
        static constexpr auto lup_decomp = gmb_lup_decomp_##S;                                              \
        static constexpr auto lu_backsolve = gmb_lu_backsolve_##S;                                          \
        static constexpr auto plu_backsolve = gmb_plu_backsolve_##S;                                        \
        static constexpr auto apply_permutation = gmb_apply_permutation_##S;                                \
        static constexpr auto plu_solve = gmb_plu_solve_##S;                                                \
        static constexpr auto linear_solve_vec = gmb_linear_solve_vec_##S;                                  \
        static constexpr auto cholesky = gmb_cholesky_##S;                                                  \
        static constexpr auto cholesky_solve = gmb_cholesky_solve_##S;                                      \
        static constexpr auto inv = gmb_inv_##S;                                                            \
        static constexpr auto host_plu_solve = gmb_host_plu_solve_##S;                                      \
        static constexpr auto host_inv = gmb_host_inv_##S;                                                  \
        static constexpr auto host_cholesky_solve = gmb_host_cholesky_solve_##S;                            \
        static constexpr auto host_plu_decomp = gmb_host_plu_decomp_##S;                                    \
        static constexpr auto orthogonal = gmb_orthogonal_##S;                                              \
        static constexpr auto nullspace = gmb_nullspace_##S;                                                \
        static constexpr auto simplex_contains = gmb_simplex_contains_##S;                                  \
        static constexpr auto trace_simplex = gmb_trace_simplex_##S;                                        \
        static constexpr auto simplex_intersect = gmb_simplex_intersect_##S;                                \
        static constexpr auto mul = gmb_mul_##S;                                                            \
        static constexpr auto xf_from_matrix = gmb_xf_from_matrix_##S;                                      \
        static constexpr auto xf_translation = gmb_xf_translation_##S;                                      \
        static constexpr auto xf_scale = gmb_xf_scale_##S;                                                  \
        static constexpr auto xf_compose = gmb_xf_compose_##S;                                              \
        static constexpr auto xf_apply = gmb_xf_apply_##S;                                                  \
    };
GMB_ABI(float, f32)
GMB_ABI(double, f64)
#undef GMB_ABI

// ---------------------------------------------------------------- device memory (RAII over the C ABI)
template <typename T>
class DeviceArray {
public:
    DeviceArray() = default;
    explicit DeviceArray(std::size_t n) : n_(n) {
        p_ = static_cast<T*>(gmb_device_alloc(n * sizeof(T)));
        if (!p_) throw Error(GMB_ERR_ALLOC, "gmb_device_alloc");
    }
    DeviceArray(DeviceArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DeviceArray& operator=(DeviceArray&& o) noexcept {
        if (this != &o) { gmb_device_free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    ~DeviceArray() { gmb_device_free(p_); }
    T* get() const { return p_; }
    std::size_t size() const { return n_; }
    void upload(const T* host, std::size_t n, void* stream = nullptr) { check(gmb_memcpy_h2d(p_, host, n * sizeof(T), stream), "h2d"); }
    void download(T* host, std::size_t n, void* stream = nullptr) const { check(gmb_memcpy_d2h(host, p_, n * sizeof(T), stream), "d2h"); }
private:
    T* p_ = nullptr;
    std::size_t n_ = 0;
};

// A batch of B records of E elements resident in HBM, in either layout.
template <typename T>
class DeviceBatch {
public:
    DeviceBatch(index_t B, int E, Layout layout)
        : B_(B), E_(E), layout_(layout),
          data_(layout == Layout::SoA ? (std::size_t)gmb_soa_elems(B, E, (int)sizeof(T)) : (std::size_t)(B * E)) {}
    index_t size() const { return B_; }
    int elems() const { return E_; }
    Layout layout() const { return layout_; }
    T* data() const { return data_.get(); }
    // AoS host array <-> this batch (coalesced AoS->SoA ingest kernel when the layout is SoA)
    void upload_aos(const T* host, void* stream = nullptr) {
        if (layout_ == Layout::AoS) { data_.upload(host, (std::size_t)(B_ * E_), stream); return; }
        DeviceArray<T> tmp((std::size_t)(B_ * E_));
        tmp.upload(host, (std::size_t)(B_ * E_), stream);
        check(Abi<T>::aos_to_soa(E_, B_, tmp.get(), data_.get(), stream), "aos_to_soa");
        check(gmb_stream_synchronize(stream), "sync");
    }
    void download_aos(T* host, void* stream = nullptr) const {
        if (layout_ == Layout::AoS) { data_.download(host, (std::size_t)(B_ * E_), stream); check(gmb_stream_synchronize(stream), "sync"); return; }
        DeviceArray<T> tmp((std::size_t)(B_ * E_));
        check(Abi<T>::soa_to_aos(E_, B_, data_.get(), tmp.get(), stream), "soa_to_aos");
        tmp.download(host, (std::size_t)(B_ * E_), stream);
        check(gmb_stream_synchronize(stream), "sync");
    }
private:
    index_t B_;
    int E_;
    Layout layout_;
    DeviceArray<T> data_;
};

// The SoA-interleaved batch container: tiles of W = 512/sizeof(T) systems, element e of system s at
// data[((s / W) * E + e) * W + s % W].
template <typename T>
class BatchSoA : public DeviceBatch<T> {
public:
    static constexpr int W = 512 / (int)sizeof(T);
    BatchSoA(index_t B, int E) : DeviceBatch<T>(B, E, Layout::SoA) {}
    static std::size_t index(index_t s, int e, int E) { return (std::size_t)(((s / W) * E + e) * W + s % W); }
};

// ================================================================= (1) device forms (asynchronous)
// Results per system: ok / parity / degenerate are uint8[B] device arrays, reorder is uint8[B][rows].

// decomp_plu<T,RowMajor>(m, rows, cols, reorder, parity) -> degenerate_ct      LUDecomp.h:188
template <typename T, bool RowMajor = true>
void decomp_plu(DeviceBatch<T>& m, index_t rows, index_t cols, std::uint8_t* reorder, std::uint8_t* parity,
                std::uint8_t* degenerate_ct, void* stream = nullptr) {
    check(Abi<T>::plu_decomp((int)rows, (int)cols, m.size(), m.data(), reorder, parity, degenerate_ct, (int)m.layout(),
                             RowMajor ? 1 : 0, stream), "decomp_plu");
}
// decomp_lup                                                                     LUDecomp.h:102
template <typename T, bool RowMajor = true>
void decomp_lup(DeviceBatch<T>& m, index_t rows, index_t cols, std::uint8_t* reorder, std::uint8_t* parity,
                std::uint8_t* degenerate_ct, void* stream = nullptr) {
    check(Abi<T>::lup_decomp((int)rows, (int)cols, m.size(), m.data(), reorder, parity, degenerate_ct, (int)m.layout(),
                             RowMajor ? 1 : 0, stream), "decomp_lup");
}
// backsolve_lu<T,RowMajor>(lu, n, m, x, skip)                                    LUDecomp.h:332
template <typename T, bool RowMajor = true>
void backsolve_lu(const DeviceBatch<T>& lu, index_t n, index_t m, DeviceBatch<T>& x, index_t skip = 0, void* stream = nullptr) {
    check(Abi<T>::lu_backsolve((int)n, (int)m, lu.size(), lu.data(), x.data(), (int)skip, (int)lu.layout(), RowMajor ? 1 : 0,
                               stream), "backsolve_lu");
}
// backsolve_plu<T,RowMajor>(plu, p, n, m, x, b, skip)                            LUDecomp.h:271
template <typename T, bool RowMajor = true>
void backsolve_plu(const DeviceBatch<T>& plu, const std::uint8_t* p, index_t n, index_t m, DeviceBatch<T>& x,
                   const DeviceBatch<T>& b, index_t skip = 0, void* stream = nullptr) {
    check(Abi<T>::plu_backsolve((int)n, (int)m, plu.size(), plu.data(), p, x.data(), b.data(), (int)skip, (int)plu.layout(),
                                RowMajor ? 1 : 0, stream), "backsolve_plu");
}
// linear_solve<T,RowMajor>(a, n, m, x, skip) -> bool                              LUDecomp.h:388
// x holds b on entry and the solution on exit (untouched where ok == 0).  `a` is NOT destroyed
// unless lu_out aliases it.
template <typename T, bool RowMajor = true>
void linear_solve(const DeviceBatch<T>& a, index_t n, index_t m, DeviceBatch<T>& x, std::uint8_t* ok, index_t skip = 0,
                  void* stream = nullptr, DeviceBatch<T>* lu_out = nullptr) {
    check(Abi<T>::plu_solve((int)n, (int)m, a.size(), a.data(), x.data(), x.data(), ok, lu_out ? lu_out->data() : nullptr,
                            (int)skip, (int)a.layout(), RowMajor ? 1 : 0, stream), "linear_solve");
}
// cholesky<T,RowMajor>(m, n) -> bool                                              Cholesky.h:37
template <typename T, bool RowMajor = true>
void cholesky(DeviceBatch<T>& m, index_t n, std::uint8_t* ok, void* stream = nullptr) {
    check(Abi<T>::cholesky((int)n, m.size(), m.data(), ok, (int)m.layout(), RowMajor ? 1 : 0, stream), "cholesky");
}
// fused Cholesky factor + solve (new)
template <typename T, bool RowMajor = true>
void cholesky_solve(const DeviceBatch<T>& a, index_t n, index_t m, DeviceBatch<T>& x, std::uint8_t* ok, void* stream = nullptr) {
    check(Abi<T>::cholesky_solve((int)n, (int)m, a.size(), a.data(), x.data(), x.data(), ok, nullptr, (int)a.layout(),
                                 RowMajor ? 1 : 0, stream), "cholesky_solve");
}
// inv(into, src) -> bool                                                          Matrix.h:717
template <typename T>
void inv(DeviceBatch<T>& into, const DeviceBatch<T>& src, index_t n, std::uint8_t* ok, bool src_row_major = true,
         bool dst_row_major = true, void* stream = nullptr) {
    check(Abi<T>::inv((int)n, src.size(), src.data(), into.data(), ok, (int)src.layout(), src_row_major ? 1 : 0,
                      dst_row_major ? 1 : 0, stream), "inv");
}

// ================================================================= (2) host forms (synchronous)
// Raw host arrays exactly as a loop over the reference's pointer functions sees them.
// `ok` receives the reference's bool per system (nullptr allowed).  Returns the number of systems
// for which the reference would have returned false.
namespace detail {
inline index_t count_false(const std::uint8_t* ok, index_t B) {
    // flags are 0 / 1 bytes: eight per step (a batch is millions of systems; a byte-wise loop showed up in the object forms)
    index_t ones = 0, i = 0;
    for (; i + 8 <= B; i += 8) {
        std::uint64_t w;
        std::memcpy(&w, ok + i, 8);
        ones += (index_t)__builtin_popcountll(w & 0x0101010101010101ull);
    }
    for (; i < B; ++i) ones += ok[i] ? 1 : 0;
    return B - ones;
}
// Makes `device` current for a scope and restores the caller's device on every exit path (exceptions included).
class DeviceScope {
public:
    explicit DeviceScope(int device) {
        if (gmb_get_device(&prev_) != GMB_OK) prev_ = -1;
        check(gmb_set_device(device), "set_device");
    }
    ~DeviceScope() {
        if (prev_ >= 0) gmb_set_device(prev_);
    }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
private:
    int prev_ = -1;
};
// The reference's bool results: a bool is one byte holding 0 or 1, which is exactly what the kernels write, so a
// caller's bool[B] receives the flags directly; a scratch array stands in when the caller passes nullptr.
static_assert(sizeof(bool) == 1, "bool[B] is used as the uint8 flag array of the C ABI");
struct OkFlags {
    std::vector<std::uint8_t> scratch;
    std::uint8_t* p;
    OkFlags(bool* ok, index_t B) : scratch(ok ? 0 : (std::size_t)B), p(ok ? reinterpret_cast<std::uint8_t*>(ok) : scratch.data()) {}
};
}  // namespace detail

// Pins an application-owned host range for the lifetime of the object (gmb_host_register): the host and object forms
// then move it over PCIe at full speed with no staging copy.  Registration failure is not an error (`pinned()` tells).
class PinnedScope {
public:
    PinnedScope(const void* p, std::size_t bytes) : p_(const_cast<void*>(p)) {
        ok_ = p_ && bytes && gmb_host_register(p_, bytes) == GMB_OK;
    }
    ~PinnedScope() {
        if (ok_) gmb_host_unregister(p_);
    }
    PinnedScope(const PinnedScope&) = delete;
    PinnedScope& operator=(const PinnedScope&) = delete;
    bool pinned() const { return ok_; }
private:
    void* p_;
    bool ok_ = false;
};

// for (s < B) ok[s] = geom::linear_solve<T,RowMajor>(a + s*n*n, n, m, x + s*n*m, skip);
// Difference from the reference: `a` is left intact (the reference leaves the LU factors in it).
template <typename T, bool RowMajor = true>
index_t linear_solve(const T* a, index_t n, index_t m, T* x, index_t B, bool* ok = nullptr, index_t skip = 0, int device = 0) {
    detail::OkFlags okb(ok, B);
    check(Abi<T>::host_plu_solve((int)n, (int)m, B, a, x, x, okb.p, (int)skip, RowMajor ? 1 : 0, device), "linear_solve");
    return detail::count_false(okb.p, B);
}

// for (s < B) degenerate[s] = geom::decomp_plu<T,RowMajor>(m + s*rows*cols, rows, cols, reorder + s*rows, parity + s);
template <typename T, bool RowMajor = true>
void decomp_plu(T* m, index_t rows, index_t cols, index_t B, index_t* reorder, bool* parity, index_t* degenerate_ct, int device = 0) {
    std::vector<std::uint8_t> p((std::size_t)(B * rows)), par((std::size_t)B), deg((std::size_t)B);
    check(Abi<T>::host_plu_decomp((int)rows, (int)cols, B, m, p.data(), par.data(), deg.data(), RowMajor ? 1 : 0, device),
          "decomp_plu");
    if (reorder) for (index_t i = 0; i < B * rows; ++i) reorder[i] = (index_t)p[(std::size_t)i];
    if (parity) for (index_t i = 0; i < B; ++i) parity[i] = par[(std::size_t)i] != 0;
    if (degenerate_ct) for (index_t i = 0; i < B; ++i) degenerate_ct[i] = (index_t)deg[(std::size_t)i];
}

// for (s < B) ok[s] = geom::cholesky<T,RowMajor>(...) followed by the L L^T solve of x (new).
template <typename T, bool RowMajor = true>
index_t cholesky_solve(const T* a, index_t n, index_t m, T* x, index_t B, bool* ok = nullptr, int device = 0) {
    detail::OkFlags okb(ok, B);
    check(Abi<T>::host_cholesky_solve((int)n, (int)m, B, a, x, x, okb.p, RowMajor ? 1 : 0, device), "cholesky_solve");
    return detail::count_false(okb.p, B);
}

// for (s < B) ok[s] = inv2x2/3x3/4x4/invNxN on flat arrays; `out` untouched where ok == false.
template <typename T>
index_t inv(T* out, const T* src, index_t n, index_t B, bool* ok = nullptr, bool src_row_major = true,
            bool dst_row_major = true, int device = 0) {
    detail::OkFlags okb(ok, B);
    check(Abi<T>::host_inv((int)n, B, src, out, okb.p, src_row_major ? 1 : 0, dst_row_major ? 1 : 0, device), "inv");
    return detail::count_false(okb.p, B);
}

// ================================================================= (3) object forms
// Arrays of the reference's own objects, duck-typed: Mx needs `elem_t`, static `ROWDIM`, `COLDIM`,
// `Layout` (0 = ROW_MAJOR, 1 = COL_MAJOR: LinalgTypes.h:36-39) and `data_begin()`; Vec needs
// `begin()`.  An array of objects that ARE their elements (sizeof(Mx) == E * sizeof(T), data_begin() == this:
// SimpleMatrix<float,4,4> is 64 packed bytes) goes to the host entry points AS IT IS -- no repacking, no staging copy;
// only objects with alignment padding (SimpleMatrix<float,3,3> is 40 bytes) are packed into a temporary.
namespace detail {
template <typename Mx> using elem_of = typename Mx::elem_t;
template <typename Mx, int E>
bool is_packed_array(const Mx* m, index_t B) {
    if (sizeof(Mx) != sizeof(elem_of<Mx>) * (std::size_t)E) return false;
    return B <= 0 || reinterpret_cast<const char*>(m[0].data_begin()) == reinterpret_cast<const char*>(&m[0]);
}
// Read-only / read-write view of an object array as the flat T[B][E] the C ABI takes.
template <typename Mx, int E>
struct FlatIn {
    std::vector<elem_of<Mx>> tmp;
    const elem_of<Mx>* p;
    FlatIn(const Mx* m, index_t B) {
        if (is_packed_array<Mx, E>(m, B)) {
            p = reinterpret_cast<const elem_of<Mx>*>(m);
        } else {
            tmp.resize((std::size_t)(B * E));
            for (index_t s = 0; s < B; ++s) std::memcpy(tmp.data() + s * E, m[s].data_begin(), sizeof(elem_of<Mx>) * (std::size_t)E);
            p = tmp.data();
        }
    }
};
template <typename Mx, int E>
struct FlatOut {
    Mx* m;
    index_t B;
    bool direct;
    std::vector<elem_of<Mx>> tmp;
    elem_of<Mx>* p;
    // load_current: the objects' present contents matter (accumulation, "untouched on failure")
    FlatOut(Mx* m_, index_t B_, bool load_current) : m(m_), B(B_), direct(is_packed_array<Mx, E>(m_, B_)) {
        if (direct) {
            p = reinterpret_cast<elem_of<Mx>*>(m);
        } else {
            tmp.resize((std::size_t)(B * E));
            if (load_current)
                for (index_t s = 0; s < B; ++s) std::memcpy(tmp.data() + s * E, m[s].data_begin(), sizeof(elem_of<Mx>) * (std::size_t)E);
            p = tmp.data();
        }
    }
    void commit(const std::uint8_t* only_where = nullptr) {
        if (direct) return;
        for (index_t s = 0; s < B; ++s)
            if (!only_where || only_where[s]) std::memcpy(m[s].data_begin(), tmp.data() + s * E, sizeof(elem_of<Mx>) * (std::size_t)E);
    }
};
}  // namespace detail

// for (s < B) ok[s] = geom::inv(&into[s], src[s]);          Matrix.h:717 (static N = 2..16)
template <typename Md, typename Mx>
index_t inv(Md* into, const Mx* src, index_t B, bool* ok = nullptr, int device = 0) {
    using T = detail::elem_of<Mx>;
    constexpr int N = (int)(Mx::ROWDIM > Mx::COLDIM ? Mx::ROWDIM : Mx::COLDIM);
    static_assert(N >= 2 && N <= GMB_MAX_N, "static dimension 2..16 required");
    static_assert(std::is_same<T, detail::elem_of<Md>>::value, "element types must agree");
    detail::FlatIn<Mx, N * N> a(src, B);
    detail::FlatOut<Md, N * N> o(into, B, /*load_current=*/true);      // current contents: "untouched on failure"
    detail::OkFlags okb(ok, B);
    check(Abi<T>::host_inv(N, B, a.p, o.p, okb.p, (int)Mx::Layout == 0 ? 1 : 0, (int)Md::Layout == 0 ? 1 : 0, device), "inv");
    o.commit(okb.p);
    return detail::count_false(okb.p, B);
}

// for (s < B) ok[s] = geom::cholesky(&m[s]);                Cholesky.h:83 (row-major arithmetic on begin())
template <typename Mx>
index_t cholesky(Mx* m, index_t B, bool* ok = nullptr, int device = 0) {
    using T = detail::elem_of<Mx>;
    constexpr int N = (int)Mx::ROWDIM;
    static_assert(Mx::ROWDIM == Mx::COLDIM && N >= 2 && N <= GMB_MAX_N, "square static dimension 2..16 required");
    // Cholesky.h:87 runs cholesky<T,true> on m->begin(), the ROW-MAJOR iterator: logical (r,c) order
    // whatever the storage layout; a symmetric input makes the two coincide for reading, and the
    // factor is written to logical (r,c).
    const bool rm = (int)Mx::Layout == 0;
    detail::DeviceScope scope(device);                       // before any allocation: the buffers must live on `device`
    detail::FlatOut<Mx, N * N> a(m, B, /*load_current=*/true);
    detail::OkFlags okb(ok, B);
    DeviceBatch<T> d(B, N * N, Layout::AoS);
    DeviceArray<std::uint8_t> dok((std::size_t)B);
    d.upload_aos(a.p);
    check(Abi<T>::cholesky(N, B, d.data(), dok.get(), GMB_LAYOUT_AOS, rm ? 1 : 0, nullptr), "cholesky");
    d.download_aos(a.p);
    dok.download(okb.p, (std::size_t)B);
    check(gmb_stream_synchronize(nullptr), "sync");
    a.commit();
    return detail::count_false(okb.p, B);
}

// for (s < B) ok[s] = geom::linear_solve<T,N>(bases[s], 1, &x[s], skip);      LUDecomp.h:417
// bases: B arrays of N column vectors (Vec<T,N>[N], contiguous N*N); x: B vectors (b in, x out).
// N < 5 is inverse-then-multiply exactly like the reference; `bases` is left intact.
template <typename V>
index_t linear_solve_vec(const V* bases, V* x, index_t N, index_t B, bool* ok = nullptr, index_t skip = 0, int device = 0) {
    using T = typename std::remove_cv<typename std::remove_reference<decltype(*std::declval<V&>().begin())>::type>::type;
    static_assert(sizeof(V) % sizeof(T) == 0, "packed vectors expected (VecBase.h:112)");
    detail::DeviceScope scope(device);
    DeviceBatch<T> dA(B, (int)(N * N), Layout::AoS), dX(B, (int)N, Layout::AoS);
    DeviceArray<std::uint8_t> dok((std::size_t)B);
    dA.upload_aos(reinterpret_cast<const T*>(bases));
    dX.upload_aos(reinterpret_cast<const T*>(x));
    check(Abi<T>::linear_solve_vec((int)N, 1, B, dA.data(), dX.data(), dX.data(), dok.get(), (int)skip, GMB_LAYOUT_AOS, nullptr),
          "linear_solve(Vec)");
    std::vector<std::uint8_t> okb((std::size_t)B);
    dX.download_aos(reinterpret_cast<T*>(x));
    dok.download(okb.data(), (std::size_t)B);
    check(gmb_stream_synchronize(nullptr), "sync");
    if (ok) for (index_t i = 0; i < B; ++i) ok[i] = okb[(std::size_t)i] != 0;
    return detail::count_false(okb.data(), B);
}

// ---- row (f)1 of the scope table: orthogonal / nullspace (Orthogonal.h:73-178) --------------------
namespace detail {
template <typename V>
using vec_elem = typename std::remove_cv<typename std::remove_reference<decltype(*std::declval<V&>().begin())>::type>::type;
}

// for (s < B) out[s] = geom::orthogonal(v + s*(N-1));        v: B groups of N-1 packed Vec<T,N>
template <typename V>
void orthogonal(const V* v, V* out, index_t N, index_t B, int device = 0) {
    using T = detail::vec_elem<V>;
    detail::DeviceScope scope(device);
    DeviceBatch<T> dV(B, (int)((N - 1) * N), Layout::AoS), dO(B, (int)N, Layout::AoS);
    dV.upload_aos(reinterpret_cast<const T*>(v));
    check(Abi<T>::orthogonal((int)N, B, dV.data(), dO.data(), GMB_LAYOUT_AOS, nullptr), "orthogonal");
    dO.download_aos(reinterpret_cast<T*>(out));
}

// for (s < B) ok[s] = geom::nullspace<T,N>(bases + s*n, n, null_basis + s*(N-n));      1 <= n < N
template <typename V>
index_t nullspace(const V* bases, index_t n, V* null_basis, index_t N, index_t B, bool* ok = nullptr, int device = 0) {
    using T = detail::vec_elem<V>;
    detail::DeviceScope scope(device);
    DeviceBatch<T> dB(B, (int)(n * N), Layout::AoS), dN(B, (int)((N - n) * N), Layout::AoS);
    DeviceArray<std::uint8_t> dok((std::size_t)B);
    dB.upload_aos(reinterpret_cast<const T*>(bases));
    check(Abi<T>::nullspace((int)N, (int)n, B, dB.data(), dN.data(), dok.get(), GMB_LAYOUT_AOS, nullptr), "nullspace");
    std::vector<std::uint8_t> okb((std::size_t)B);
    dN.download_aos(reinterpret_cast<T*>(null_basis));
    dok.download(okb.data(), (std::size_t)B);
    check(gmb_stream_synchronize(nullptr), "sync");
    if (ok) for (index_t i = 0; i < B; ++i) ok[i] = okb[(std::size_t)i] != 0;
    return detail::count_false(okb.data(), B);
}

// ---- row (f)2 of the scope table: batched simplex queries (shape/Simplex.h) ------------------------
// Object forms, duck-typed on the members the reference's types have: Simplex::pts / ::n (Simplex.h:92-94),
// Ray::origin / ::direction (Ray.h:24-26), Rect<T,1>::lo / ::hi.  Object arrays are uploaded as they are
// and addressed through the strides of the C ABI; nothing is repacked on the host.
namespace detail {
template <typename S>
using simplex_elem = vec_elem<decltype(std::declval<S&>().pts[0])>;

template <typename T, typename Obj, typename Member>
inline std::size_t member_offset(const Obj* o, const Member* m, const char* what) {
    const std::size_t off = (std::size_t)(reinterpret_cast<const char*>(m) - reinterpret_cast<const char*>(o));
    if (off % sizeof(T) != 0 || sizeof(Obj) % sizeof(T) != 0) throw Error(GMB_ERR_INVALID_ARG, what);
    return off / sizeof(T);
}

// uploads `count` objects; returns the device array (as elements of T)
template <typename T, typename Obj>
inline DeviceArray<T> upload_objects(const Obj* objs, index_t count) {
    DeviceArray<T> d((std::size_t)count * sizeof(Obj) / sizeof(T) + 1);
    if (count > 0) d.upload(reinterpret_cast<const T*>(objs), (std::size_t)count * sizeof(Obj) / sizeof(T));
    return d;
}

template <typename S>
inline DeviceArray<std::uint8_t> upload_vertex_counts(const S* simplexes, index_t B) {
    std::vector<std::uint8_t> nv((std::size_t)B);
    for (index_t i = 0; i < B; ++i) nv[(std::size_t)i] = (std::uint8_t)simplexes[i].n;
    DeviceArray<std::uint8_t> d((std::size_t)B + 1);
    if (B > 0) d.upload(nv.data(), (std::size_t)B);
    return d;
}
}  // namespace detail

// for (s < B) inside[s] = simplexes[s].contains(p[s]);       -> number of points inside      (Simplex.h:450)
template <typename S, typename V>
index_t contains(const S* simplexes, const V* p, index_t B, bool* inside, int device = 0) {
    using T = detail::simplex_elem<S>;
    constexpr int N = (int)(sizeof(V) / sizeof(T));
    detail::DeviceScope scope(device);
    if (B <= 0) return 0;
    const std::size_t voff = detail::member_offset<T>(simplexes, simplexes[0].pts[0].begin(), "Simplex layout");
    DeviceArray<T> dS = detail::upload_objects<T>(simplexes, B), dP = detail::upload_objects<T>(p, B);
    DeviceArray<std::uint8_t> dn = detail::upload_vertex_counts(simplexes, B), din((std::size_t)B);
    check(Abi<T>::simplex_contains(N, B, dS.get() + voff, (std::int64_t)(sizeof(S) / sizeof(T)), dn.get(), dP.get(), din.get(),
                                   GMB_LAYOUT_AOS, nullptr), "simplex_contains");
    std::vector<std::uint8_t> inb((std::size_t)B);
    din.download(inb.data(), (std::size_t)B);
    check(gmb_stream_synchronize(nullptr), "sync");
    index_t n_in = 0;
    for (index_t i = 0; i < B; ++i) {
        if (inside) inside[i] = inb[(std::size_t)i] != 0;
        n_in += inb[(std::size_t)i] != 0;
    }
    return n_in;
}

// for (s < B) hit[s] = geom::trace_simplex<T,N>(verts + s*N, rays[s], uv ? &uv[s] : nullptr, &t[s]);   (Simplex.h:647)
// verts: B groups of N packed Vec<T,N>; t[s] and uv[s] are written only where hit.  -> number of hits.
// t[s] is the coefficient of -direction, x[N-1]; the reference itself stores x[N] (out of bounds, :675).
template <typename V, typename R, typename VU, typename T>
index_t trace_simplex(const V* verts, const R* rays, VU* uv, T* t, index_t B, bool* hit, int device = 0) {
    static_assert(std::is_same<T, detail::vec_elem<V>>::value, "ray parameter type must be the vertices' element type");
    constexpr int N = (int)(sizeof(V) / sizeof(T));
    detail::DeviceScope scope(device);
    if (B <= 0) return 0;
    const std::size_t ooff = detail::member_offset<T>(rays, rays[0].origin.begin(), "Ray layout");
    const std::size_t doff = detail::member_offset<T>(rays, rays[0].direction.begin(), "Ray layout");
    DeviceArray<T> dV = detail::upload_objects<T>(verts, B * N), dR = detail::upload_objects<T>(rays, B);
    DeviceArray<T> dt((std::size_t)B), duv(uv ? (std::size_t)B * (N - 1) : 1);
    DeviceArray<std::uint8_t> dh((std::size_t)B);
    check(Abi<T>::trace_simplex(N, B, dV.get(), 0, dR.get() + ooff, dR.get() + doff, (std::int64_t)(sizeof(R) / sizeof(T)), dh.get(),
                                dt.get(), uv ? duv.get() : nullptr, GMB_LAYOUT_AOS, nullptr), "trace_simplex");
    std::vector<std::uint8_t> hb((std::size_t)B);
    std::vector<T> tb((std::size_t)B), uvb(uv ? (std::size_t)B * (N - 1) : 0);
    dh.download(hb.data(), (std::size_t)B);
    dt.download(tb.data(), (std::size_t)B);
    if (uv) duv.download(uvb.data(), uvb.size());
    check(gmb_stream_synchronize(nullptr), "sync");
    index_t hits = 0;
    for (index_t i = 0; i < B; ++i) {
        const bool h = hb[(std::size_t)i] != 0;
        if (hit) hit[i] = h;
        hits += h;
        if (!h) continue;
        t[i] = tb[(std::size_t)i];
        if (uv) std::memcpy(reinterpret_cast<T*>(&uv[i]), uvb.data() + (std::size_t)i * (N - 1), sizeof(T) * (N - 1));
    }
    return hits;
}

// for (s < B) intervals[s] = simplexes[s].intersect(rays[s]);   -> number of non-empty intervals   (Simplex.h:251)
template <typename S, typename R, typename I>
index_t intersect(const S* simplexes, const R* rays, index_t B, I* intervals, int device = 0) {
    using T = detail::simplex_elem<S>;
    constexpr int N = (int)(sizeof(R) / sizeof(T) / 2);
    detail::DeviceScope scope(device);
    if (B <= 0) return 0;
    const std::size_t voff = detail::member_offset<T>(simplexes, simplexes[0].pts[0].begin(), "Simplex layout");
    const std::size_t ooff = detail::member_offset<T>(rays, rays[0].origin.begin(), "Ray layout");
    const std::size_t doff = detail::member_offset<T>(rays, rays[0].direction.begin(), "Ray layout");
    DeviceArray<T> dS = detail::upload_objects<T>(simplexes, B), dR = detail::upload_objects<T>(rays, B), dI((std::size_t)B * 2);
    DeviceArray<std::uint8_t> dn = detail::upload_vertex_counts(simplexes, B);
    check(Abi<T>::simplex_intersect(N, B, dS.get() + voff, (std::int64_t)(sizeof(S) / sizeof(T)), dn.get(), dR.get() + ooff,
                                    dR.get() + doff, (std::int64_t)(sizeof(R) / sizeof(T)), dI.get(), GMB_LAYOUT_AOS, nullptr),
          "simplex_intersect");
    std::vector<T> ib((std::size_t)B * 2);
    dI.download(ib.data(), ib.size());
    check(gmb_stream_synchronize(nullptr), "sync");
    index_t nonempty = 0;
    for (index_t i = 0; i < B; ++i) {
        intervals[i].lo = ib[2 * (std::size_t)i];
        intervals[i].hi = ib[2 * (std::size_t)i + 1];
        nonempty += ib[2 * (std::size_t)i] <= ib[2 * (std::size_t)i + 1];
    }
    return nonempty;
}

// ---- row (f)4 of the scope table: mul / mul_acc (Matrix.h:386-456 -> MatrixMult.h:54-80) ------------
// for (s < B) geom::mul(&into[s], a[s], b[s]);   (accumulate: geom::mul_acc)      `into` may be `a` or `b`
// for shapes up to 4 x 4 x 4.  Matrix objects as in the other object forms (elem_t, ROWDIM, COLDIM, Layout,
// data_begin()); dimensions 1..16.
template <typename Md, typename Ma, typename Mb>
void mul(Md* into, const Ma* a, const Mb* b, index_t B, bool accumulate = false, int device = 0) {
    using T = detail::elem_of<Ma>;
    constexpr int R = (int)Ma::ROWDIM, K = (int)Ma::COLDIM, C = (int)Mb::COLDIM;
    static_assert((int)Mb::ROWDIM == K && (int)Md::ROWDIM == R && (int)Md::COLDIM == C, "dimension mismatch");
    static_assert(std::is_same<T, detail::elem_of<Mb>>::value && std::is_same<T, detail::elem_of<Md>>::value, "element types must agree");
    detail::DeviceScope scope(device);
    if (B <= 0) return;
    detail::FlatIn<Ma, R * K> ha(a, B);
    detail::FlatIn<Mb, K * C> hb(b, B);
    detail::FlatOut<Md, R * C> hd(into, B, /*load_current=*/accumulate);
    DeviceBatch<T> dA(B, R * K, Layout::AoS), dB(B, K * C, Layout::AoS), dD(B, R * C, Layout::AoS);
    dA.upload_aos(ha.p);
    dB.upload_aos(hb.p);
    if (accumulate) dD.upload_aos(hd.p);
    check(Abi<T>::mul(R, K, C, B, dA.data(), dB.data(), dD.data(), accumulate ? 1 : 0, GMB_LAYOUT_AOS, (int)Ma::Layout == 0 ? 1 : 0,
                      (int)Mb::Layout == 0 ? 1 : 0, (int)Md::Layout == 0 ? 1 : 0, nullptr), "mul");
    dD.download_aos(hd.p);
    check(gmb_stream_synchronize(nullptr), "sync");
    hd.commit();
}
template <typename Md, typename Ma, typename Mb>
void mul_acc(Md* into, const Ma* a, const Mb* b, index_t B, int device = 0) { mul(into, a, b, B, true, device); }

// ---- row (f)3 of the scope table: AffineTransform (linalg/AffineTransform.h) -------------------------
// Object forms, duck-typed on AffineTransform's members `mat` and `inv` (:91-93, SimpleMatrix<T,N+1,N+1>).
// The device works on packed records (mat | inv, row-major); objects are packed / unpacked here because
// SimpleMatrix can carry alignment padding.
namespace detail {
template <typename X> using xf_elem = elem_of<typename std::remove_reference<decltype(std::declval<X&>().mat)>::type>;
template <typename X> constexpr int xf_K() { return (int)std::remove_reference<decltype(std::declval<X&>().mat)>::type::ROWDIM; }

template <typename X>
std::vector<xf_elem<X>> pack_xf(const X* xf, index_t B) {
    using T = xf_elem<X>;
    constexpr int KK = xf_K<X>() * xf_K<X>();
    std::vector<T> v((std::size_t)(B * 2 * KK));
    for (index_t s = 0; s < B; ++s) {
        std::memcpy(v.data() + s * 2 * KK, xf[s].mat.data_begin(), sizeof(T) * KK);
        std::memcpy(v.data() + s * 2 * KK + KK, xf[s].inv.data_begin(), sizeof(T) * KK);
    }
    return v;
}
template <typename X>
void unpack_xf(X* xf, const std::vector<xf_elem<X>>& v, index_t B) {
    using T = xf_elem<X>;
    constexpr int KK = xf_K<X>() * xf_K<X>();
    for (index_t s = 0; s < B; ++s) {
        std::memcpy(xf[s].mat.data_begin(), v.data() + s * 2 * KK, sizeof(T) * KK);
        std::memcpy(xf[s].inv.data_begin(), v.data() + s * 2 * KK + KK, sizeof(T) * KK);
    }
}
}  // namespace detail

// for (s < B) out[s] = geom::transformation(m[s]);       AffineTransform.h:862.  -> number of singular matrices
// (whose transforms carry inv = identity, as in the reference).
template <typename X, typename Mx>
index_t transformation(X* out, const Mx* m, index_t B, int device = 0) {
    using T = detail::elem_of<Mx>;
    constexpr int N = (int)Mx::ROWDIM, E = 2 * (N + 1) * (N + 1);
    static_assert(detail::xf_K<X>() == N + 1 && (int)Mx::COLDIM == N, "AffineTransform<T,N> from an N x N matrix");
    detail::DeviceScope scope(device);
    if (B <= 0) return 0;
    detail::FlatIn<Mx, N * N> hm(m, B);
    DeviceBatch<T> dM(B, N * N, Layout::AoS), dX(B, E, Layout::AoS);
    DeviceArray<std::uint8_t> dok((std::size_t)B);
    dM.upload_aos(hm.p);
    check(Abi<T>::xf_from_matrix(N, B, dM.data(), dX.data(), dok.get(), GMB_LAYOUT_AOS, (int)Mx::Layout == 0 ? 1 : 0, nullptr),
          "transformation");
    std::vector<T> hx((std::size_t)(B * E));
    std::vector<std::uint8_t> okb((std::size_t)B);
    dX.download_aos(hx.data());
    dok.download(okb.data(), (std::size_t)B);
    check(gmb_stream_synchronize(nullptr), "sync");
    detail::unpack_xf(out, hx, B);
    return detail::count_false(okb.data(), B);
}

namespace detail {
template <typename X, typename V>
void xf_from_vectors(X* out, const V* v, index_t B, bool is_scale, int device) {
    using T = xf_elem<X>;
    constexpr int N = xf_K<X>() - 1, E = 2 * (N + 1) * (N + 1);
    static_assert(sizeof(V) == sizeof(T) * N, "packed Vec<T,N> expected");
    detail::DeviceScope scope(device);
    if (B <= 0) return;
    DeviceBatch<T> dV(B, N, Layout::AoS), dX(B, E, Layout::AoS);
    dV.upload_aos(reinterpret_cast<const T*>(v));
    check((is_scale ? Abi<T>::xf_scale : Abi<T>::xf_translation)(N, B, dV.data(), dX.data(), GMB_LAYOUT_AOS, nullptr), "translation/scale");
    std::vector<T> hx((std::size_t)(B * E));
    dX.download_aos(hx.data());
    check(gmb_stream_synchronize(nullptr), "sync");
    unpack_xf(out, hx, B);
}
}  // namespace detail

// for (s < B) out[s] = geom::translation(tx[s]);  /  geom::scale(sx[s]);        AffineTransform.h:814 / :838
template <typename X, typename V>
void translation(X* out, const V* tx, index_t B, int device = 0) { detail::xf_from_vectors(out, tx, B, false, device); }
template <typename X, typename V>
void scale(X* out, const V* sx, index_t B, int device = 0) { detail::xf_from_vectors(out, sx, B, true, device); }

// for (s < B) out[s] = xf1[s] * xf2[s];   (inverse: xf1[s] / xf2[s])            AffineTransform.h:142 / :154
// `out` may be xf1 or xf2 (the host copies are independent of the inputs).
template <typename X>
void compose(X* out, const X* xf1, const X* xf2, index_t B, bool inverse = false, int device = 0) {
    using T = detail::xf_elem<X>;
    constexpr int N = detail::xf_K<X>() - 1, E = 2 * (N + 1) * (N + 1);
    detail::DeviceScope scope(device);
    if (B <= 0) return;
    auto h1 = detail::pack_xf(xf1, B), h2 = detail::pack_xf(xf2, B);
    DeviceBatch<T> d1(B, E, Layout::AoS), d2(B, E, Layout::AoS), dO(B, E, Layout::AoS);
    d1.upload_aos(h1.data());
    d2.upload_aos(h2.data());
    check(Abi<T>::xf_compose(N, B, d1.data(), d2.data(), dO.data(), inverse ? 1 : 0, GMB_LAYOUT_AOS, nullptr), "compose");
    dO.download_aos(h1.data());
    check(gmb_stream_synchronize(nullptr), "sync");
    detail::unpack_xf(out, h1, B);
}

// The apply family on point arrays (AffineTransform.h:224-293); mode = GMB_XF_APPLY ... GMB_XF_APPLY_INVERSE_NORMAL.
//   one transform for the whole array:   for (s < B) out[s] = xf.apply(p[s]);
template <typename X, typename V, typename = decltype(std::declval<const X&>().mat)>
void apply(const X& xf, const V* p, V* out, index_t B, int mode = GMB_XF_APPLY, int device = 0) {
    using T = detail::xf_elem<X>;
    constexpr int N = detail::xf_K<X>() - 1, E = 2 * (N + 1) * (N + 1);
    static_assert(sizeof(V) == sizeof(T) * N, "packed Vec<T,N> expected");
    detail::DeviceScope scope(device);
    if (B <= 0) return;
    auto hx = detail::pack_xf(&xf, 1);
    DeviceArray<T> dX((std::size_t)E);
    DeviceBatch<T> dP(B, N, Layout::AoS), dO(B, N, Layout::AoS);
    dX.upload(hx.data(), (std::size_t)E);
    dP.upload_aos(reinterpret_cast<const T*>(p));
    check(Abi<T>::xf_apply(N, mode, B, dX.get(), 1, dP.data(), dO.data(), GMB_LAYOUT_AOS, nullptr), "apply");
    dO.download_aos(reinterpret_cast<T*>(out));
    check(gmb_stream_synchronize(nullptr), "sync");
}
//   one transform per point:             for (s < B) out[s] = xf[s].apply(p[s]);
template <typename X, typename V>
void apply(const X* xf, const V* p, V* out, index_t B, int mode = GMB_XF_APPLY, int device = 0) {
    using T = detail::xf_elem<X>;
    constexpr int N = detail::xf_K<X>() - 1, E = 2 * (N + 1) * (N + 1);
    static_assert(sizeof(V) == sizeof(T) * N, "packed Vec<T,N> expected");
    detail::DeviceScope scope(device);
    if (B <= 0) return;
    auto hx = detail::pack_xf(xf, B);
    DeviceBatch<T> dX(B, E, Layout::AoS), dP(B, N, Layout::AoS), dO(B, N, Layout::AoS);
    dX.upload_aos(hx.data());
    dP.upload_aos(reinterpret_cast<const T*>(p));
    check(Abi<T>::xf_apply(N, mode, B, dX.data(), B, dP.data(), dO.data(), GMB_LAYOUT_AOS, nullptr), "apply");
    dO.download_aos(reinterpret_cast<T*>(out));
    check(gmb_stream_synchronize(nullptr), "sync");
}

}  // namespace batch
}  // namespace geom
