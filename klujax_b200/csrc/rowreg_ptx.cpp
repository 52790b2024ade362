// rowreg_ptx.cpp — renders a register-resident plan (rowreg.hpp) as straight-line PTX for sm_100a.
//
// One warp = one system; the CTA is just a bundle of independent warps.  Everything that depends
// on the sparsity pattern is an immediate of the generated code: register names, source lanes of
// the shuffles, lane masks.  Per-lane shared-memory addresses come from 16-bit tables (global
// memory, identical for every system, L1/L2 resident).
//
// Kernel parameters (all device pointers):
//   Ax [L][ax_stride] values in the order of THIS call's COO list, b [L][b_stride], x [L][n*r],
//   tabs (uint16), dst (int32 [nz]: CSC position of each COO entry of the call, -1 = outside the
//   analysed pattern), status (int32 [L]), growth (double [L]), bad (int32: pattern flag),
//   ax_stride, b_stride (elements), L.
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "rowreg.hpp"

namespace kb {
namespace rr {
namespace {

struct W {
    std::string s;
    void operator()(const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        s += buf;
        s += '\n';
    }
};

}  // namespace

bool emit_ptx(const Plan& p, PtxKernel& out) {
    if (!p.ok) return false;
    const bool cx = p.is_complex;
    const int VS = cx ? 16 : 8;
    const int n = p.n, r = p.r, nz = p.nz;
    // shared-memory layout of one warp: the slot pool (entries of A until they are loaded, parked
    // values afterwards), the right-hand side, x of the blocks solved so far (BTF), Rs
    size_t off = (static_cast<size_t>(std::max(1, p.planes)) * kLanes * VS + 15) & ~size_t(15);
    const size_t off_u = 0;
    const size_t off_b = off;
    off += static_cast<size_t>(n) * r * VS;
    const size_t off_x = off;
    if (p.need_xbuf) off += static_cast<size_t>(n) * r * VS;
    const size_t off_rs = off;
    off += static_cast<size_t>(n) * 8;
    off = (off + 127) & ~size_t(127);
    out.smem_per_warp = off;
    out.P = p.P;
    const int WPC = out.warps_per_cta;
    out.name = "klujax_rowreg";

    W body;
    // helpers ---------------------------------------------------------------------------------
    auto tabload = [&](const char* dst, int tab) {  // u16 table entry of this lane -> u32 dst
        body("ld.global.nc.u16 %%h0, [%%tabl+%d];", tab * kLanes * 2);
        body("cvt.u32.u16 %s, %%h0;", dst);
    };
    auto maskpred = [&](const char* pr, uint32_t mask) {
        if (mask == 0xffffffffu) {
            body("setp.eq.u32 %s, 0, 0;", pr);
        } else {
            body("and.b32 %%t0, %%lanebit, 0x%x;", mask);
            body("setp.ne.u32 %s, %%t0, 0;", pr);
        }
    };
    auto shfl64 = [&](const char* dst, const char* src, int lane) {
        body("mov.b64 {%%s0, %%s1}, %s;", src);
        body("shfl.sync.idx.b32 %%s2, %%s0, %d, 0x1f, 0xffffffff;", lane);
        body("shfl.sync.idx.b32 %%s3, %%s1, %d, 0x1f, 0xffffffff;", lane);
        body("mov.b64 %s, {%%s2, %%s3};", dst);
    };
    char rx[32], ry[32], a1[48], a2[48], a3[48], a4[48];
    const bool lockstep = out.lockstep;
    std::vector<std::string> lz_keys;  // masked multipliers of the current pivot: "reg:Ltemp:mask"
    int lz_base = 0;
    auto RX = [&](int q) { snprintf(rx, sizeof rx, "%%rx%d", q); return rx; };
    auto RY = [&](int q) { snprintf(ry, sizeof ry, "%%ry%d", q); return ry; };

    // phase marks for the in-kernel cycle counters (prof != 0): 0 staging | 1 scale + load |
    // 2 elimination | 3 back substitution + store
    int phase = 1;
    size_t last_elim = 0;
    for (size_t i = 0; i < p.ins.size(); ++i) {
        const uint8_t op = p.ins[i].op;
        if (op == OP_PIVOT || op == OP_LMUL || op == OP_DUMP || (op == OP_MAC && p.ins[i].b != -2)) last_elim = i;
    }
    // the last block's back substitution starts after its last PIVOT
    size_t last_pivot = 0;
    for (size_t i = 0; i < p.ins.size(); ++i)
        if (p.ins[i].op == OP_PIVOT || p.ins[i].op == OP_LMUL || p.ins[i].op == OP_DUMP) last_pivot = i;
    (void)last_elim;
    size_t idx = 0;
    for (const Ins& in : p.ins) {
        const size_t cur = idx++;
        if (phase == 1 && in.op == OP_PIVOT) {
            body("mov.u64 %%ck2, %%clock64;");
            phase = 2;
        }
        if (phase == 2 && cur > last_pivot && (in.op == OP_SYNC || in.op == OP_XMUL || in.op == OP_XMUL_S)) {
            body("mov.u64 %%ck3, %%clock64;");
            phase = 3;
        }
        switch (in.op) {
            case OP_ROWSCALE: {
                // lane handles original row i = q*32 + lane
                const int q = in.c, ml = in.b;
                body("// ---- row scale factors, rows %d..", q * kLanes);
                body("mov.f64 %%fc, 0d0000000000000000;");
                for (int e = 0; e < ml; ++e) {
                    tabload("%t1", in.tab + e);
                    body("setp.ne.u32 %%p1, %%t1, 65535;");
                    body("mad.lo.u32 %%t2, %%t1, %d, %%sa;", VS);
                    if (cx) {
                        body("@%%p1 ld.shared.v2.f64 {%%fa, %%fb}, [%%t2];");
                        body("@%%p1 abs.f64 %%fa, %%fa;");
                        body("@%%p1 abs.f64 %%fb, %%fb;");
                        body("@%%p1 max.f64 %%fa, %%fa, %%fb;");
                    } else {
                        body("@%%p1 ld.shared.f64 %%fa, [%%t2];");
                        body("@%%p1 abs.f64 %%fa, %%fa;");
                    }
                    body("@%%p1 max.f64 %%fc, %%fc, %%fa;");
                }
                // rs = c * sqrt(max |a/c|^2) for complex, c for real; 1 for an empty / non-finite row
                body("mov.f64 %%frs, 0d3FF0000000000000;");
                body("setp.gt.f64 %%p2, %%fc, 0d0000000000000000;");
                body("setp.lt.and.f64 %%p2, %%fc, 0d7FF0000000000000, %%p2;");
                if (cx) {
                    body("rcp.rn.f64 %%fic, %%fc;");
                    body("mov.f64 %%fm, 0d0000000000000000;");
                    for (int e = 0; e < ml; ++e) {
                        tabload("%t1", in.tab + e);
                        body("setp.ne.u32 %%p1, %%t1, 65535;");
                        body("mad.lo.u32 %%t2, %%t1, %d, %%sa;", VS);
                        body("@%%p1 ld.shared.v2.f64 {%%fa, %%fb}, [%%t2];");
                        body("@%%p1 mul.f64 %%fa, %%fa, %%fic;");
                        body("@%%p1 mul.f64 %%fb, %%fb, %%fic;");
                        body("@%%p1 mul.f64 %%fa, %%fa, %%fa;");
                        body("@%%p1 fma.rn.f64 %%fa, %%fb, %%fb, %%fa;");
                        body("@%%p1 max.f64 %%fm, %%fm, %%fa;");
                    }
                    body("sqrt.rn.f64 %%fm, %%fm;");
                    body("@%%p2 mul.f64 %%frs, %%fc, %%fm;");
                } else {
                    body("@%%p2 mov.f64 %%frs, %%fc;");
                }
                body("rcp.rn.f64 %%fic, %%frs;");
                for (int e = 0; e < ml; ++e) {
                    tabload("%t1", in.tab + e);
                    body("setp.ne.u32 %%p1, %%t1, 65535;");
                    body("mad.lo.u32 %%t2, %%t1, %d, %%sa;", VS);
                    if (cx) {
                        body("@%%p1 ld.shared.v2.f64 {%%fa, %%fb}, [%%t2];");
                        body("@%%p1 mul.f64 %%fa, %%fa, %%fic;");
                        body("@%%p1 mul.f64 %%fb, %%fb, %%fic;");
                        body("@%%p1 st.shared.v2.f64 [%%t2], {%%fa, %%fb};");
                    } else {
                        body("@%%p1 ld.shared.f64 %%fa, [%%t2];");
                        body("@%%p1 mul.f64 %%fa, %%fa, %%fic;");
                        body("@%%p1 st.shared.f64 [%%t2], %%fa;");
                    }
                }
                body("add.u32 %%t1, %%lane, %d;", q * kLanes);
                body("setp.lt.u32 %%p1, %%t1, %d;", n);
                body("mad.lo.u32 %%t2, %%t1, 8, %%srs;");
                body("@%%p1 st.shared.f64 [%%t2], %%frs;");
                break;
            }
            case OP_LOADV: {
                maskpred("%p1", in.mask);
                if (cx)
                    body("@%%p1 ld.shared.v2.f64 {%s, %s}, [%%sl+%d];", RX(in.a), RY(in.a), in.t * kLanes * VS);
                else
                    body("@%%p1 ld.shared.f64 %s, [%%sl+%d];", RX(in.a), in.t * kLanes * VS);
                break;
            }
            case OP_LOADB: {
                tabload("%t1", in.tab);
                body("setp.ne.u32 %%p1, %%t1, 65535;");
                body("mad.lo.u32 %%t2, %%t1, %d, %%sb;", VS);
                body("mov.f64 %s, 0d0000000000000000;", RX(in.a));
                if (cx) {
                    body("mov.f64 %s, 0d0000000000000000;", RY(in.a));
                    body("@%%p1 ld.shared.v2.f64 {%s, %s}, [%%t2];", RX(in.a), RY(in.a));
                } else {
                    body("@%%p1 ld.shared.f64 %s, [%%t2];", RX(in.a));
                }
                bool any = false;
                for (int l = 0; l < kLanes; ++l) any |= p.tabs[static_cast<size_t>(in.tab2) * kLanes + l] != kNone;
                if (any) {
                    tabload("%t1", in.tab2);
                    body("setp.ne.u32 %%p1, %%t1, 65535;");
                    body("mad.lo.u32 %%t2, %%t1, 8, %%srs;");
                    body("mov.f64 %%fa, 0d3FF0000000000000;");
                    body("@%%p1 ld.shared.f64 %%fa, [%%t2];");
                    body("div.rn.f64 %s, %s, %%fa;", RX(in.a), RX(in.a));
                    if (cx) body("div.rn.f64 %s, %s, %%fa;", RY(in.a), RY(in.a));
                }
                break;
            }
            case OP_OFFMAC: {
                tabload("%t1", in.tab);
                body("setp.ne.u32 %%p1, %%t1, 65535;");
                body("mad.lo.u32 %%t2, %%t1, %d, %%sa;", VS);
                tabload("%t3", in.tab2);
                body("setp.ne.and.u32 %%p1, %%t3, 65535, %%p1;");
                body("mad.lo.u32 %%t3, %%t3, %d, %%sx;", VS);
                if (cx) {
                    body("@%%p1 ld.shared.v2.f64 {%%fa, %%fb}, [%%t2];");
                    body("@%%p1 ld.shared.v2.f64 {%%fc, %%fm}, [%%t3];");
                    body("neg.f64 %%fic, %%fa;");
                    body("neg.f64 %%frs, %%fb;");
                    body("@%%p1 fma.rn.f64 %s, %%fic, %%fc, %s;", RX(in.a), RX(in.a));
                    body("@%%p1 fma.rn.f64 %s, %%fb, %%fm, %s;", RX(in.a), RX(in.a));
                    body("@%%p1 fma.rn.f64 %s, %%fic, %%fm, %s;", RY(in.a), RY(in.a));
                    body("@%%p1 fma.rn.f64 %s, %%frs, %%fc, %s;", RY(in.a), RY(in.a));
                } else {
                    body("@%%p1 ld.shared.f64 %%fa, [%%t2];");
                    body("@%%p1 ld.shared.f64 %%fc, [%%t3];");
                    body("neg.f64 %%fic, %%fa;");
                    body("@%%p1 fma.rn.f64 %s, %%fic, %%fc, %s;", RX(in.a), RX(in.a));
                }
                break;
            }
            case OP_PIVOT: {
                lz_base += static_cast<int>(lz_keys.size());
                lz_keys.clear();
                // The code is one long straight line (hundreds of KB): the warps of the CTA walk it
                // together, so that an instruction line fetched by one is in the instruction cache for
                // the others (ncu: 44 % of the stall samples were "no instruction" without this).
                if (lockstep) body("bar.sync 0;");
                body("// ---- pivot (register %d, lane %d)", in.a, in.b);
                shfl64("%dx", RX(in.a), in.b);
                if (cx) {
                    // 1/d = conj(d) / |d|^2.  Rows are scaled to max |a| = 1, so |d|^2 leaves the range
                    // of a double only for a pivot that is numerically zero (or after overflow):
                    // both are reported as KLU_SINGULAR.
                    shfl64("%dy", RY(in.a), in.b);
                    body("mul.f64 %%fm, %%dx, %%dx;");
                    body("fma.rn.f64 %%fm, %%dy, %%dy, %%fm;");
                    body("setp.gt.f64 %%p1, %%fm, 0d05E0000000000000;");           // ~1e-280
                    body("setp.lt.and.f64 %%p1, %%fm, 0d7A10000000000000, %%p1;");  // ~1e280
                    body("@!%%p1 mov.u32 %%bad, 1;");
                    body("rcp.rn.f64 %%fic, %%fm;");
                    body("mul.f64 %%ix%d, %%dx, %%fic;", in.t);
                    body("neg.f64 %%fic, %%fic;");
                    body("mul.f64 %%iy%d, %%dy, %%fic;", in.t);
                } else {
                    body("abs.f64 %%fa, %%dx;");
                    body("setp.gt.f64 %%p1, %%fa, 0d0000000000000000;");
                    body("setp.lt.and.f64 %%p1, %%fa, 0d7FF0000000000000, %%p1;");
                    body("@!%%p1 mov.u32 %%bad, 1;");
                    body("rcp.rn.f64 %%ix%d, %%dx;", in.t);
                }
                break;
            }
            case OP_STOREINV: {
                body("setp.eq.u32 %%p1, %%lane, %d;", in.b);
                body("@%%p1 mov.f64 %s, %%ix%d;", RX(in.a), in.t);
                if (cx) body("@%%p1 mov.f64 %s, %%iy%d;", RY(in.a), in.t);
                break;
            }
            case OP_LMUL: {
                maskpred("%p1", in.mask);
                if (cx) {
                    body("mul.f64 %%fa, %s, %%ix%d;", RX(in.a), in.t);
                    body("neg.f64 %%fc, %s;", RY(in.a));
                    body("fma.rn.f64 %%fa, %%fc, %%iy%d, %%fa;", in.t);
                    body("mul.f64 %%fb, %s, %%iy%d;", RX(in.a), in.t);
                    body("fma.rn.f64 %%fb, %s, %%ix%d, %%fb;", RY(in.a), in.t);
                    body("@%%p1 mov.f64 %s, %%fa;", RX(in.a));
                    body("@%%p1 mov.f64 %s, %%fb;", RY(in.a));
                    body("mul.f64 %%fm, %%fa, %%fa;");
                    body("fma.rn.f64 %%fm, %%fb, %%fb, %%fm;");
                } else {
                    body("mul.f64 %%fa, %s, %%ix%d;", RX(in.a), in.t);
                    body("@%%p1 mov.f64 %s, %%fa;", RX(in.a));
                    body("mul.f64 %%fm, %%fa, %%fa;");
                }
                // max that lets a NaN through (a NaN multiplier must fail the certificate)
                body("setp.gtu.and.f64 %%p2, %%fm, %%gmax, %%p1;");
                body("@%%p2 mov.f64 %%gmax, %%fm;");
                break;
            }
            case OP_BCAST: {
                snprintf(a1, sizeof a1, "%%ux%d", in.t);
                shfl64(a1, RX(in.a), in.b);
                if (cx) {
                    snprintf(a2, sizeof a2, "%%uy%d", in.t);
                    shfl64(a2, RY(in.a), in.b);
                }
                break;
            }
            case OP_ZERO: {
                maskpred("%p1", in.mask);
                body("@%%p1 mov.f64 %s, 0d0000000000000000;", RX(in.a));
                if (cx) body("@%%p1 mov.f64 %s, 0d0000000000000000;", RY(in.a));
                break;
            }
            case OP_LSPREAD: {
                tabload("%t1", in.tab);
                body("mov.b64 {%%s0, %%s1}, %s;", RX(in.a));
                body("shfl.sync.idx.b32 %%s2, %%s0, %%t1, 0x1f, 0xffffffff;");
                body("shfl.sync.idx.b32 %%s3, %%s1, %%t1, 0x1f, 0xffffffff;");
                body("mov.b64 %%lx%d, {%%s2, %%s3};", in.t);
                if (cx) {
                    body("mov.b64 {%%s0, %%s1}, %s;", RY(in.a));
                    body("shfl.sync.idx.b32 %%s2, %%s0, %%t1, 0x1f, 0xffffffff;");
                    body("shfl.sync.idx.b32 %%s3, %%s1, %%t1, 0x1f, 0xffffffff;");
                    body("mov.b64 %%ly%d, {%%s2, %%s3};", in.t);
                }
                break;
            }
            case OP_MAC: {
                // Lanes outside the mask hold other entries in R[a]: they multiply by ZERO instead of
                // being predicated off (the assembler turns a predicated FMA into FMA + two selects;
                // a masked multiplier costs its selects once per pivot and mask, not per multiply-add).
                char lkey[64];
                snprintf(lkey, sizeof lkey, "%d:%d:%u", in.b, in.b >= 0 ? 0 : in.t2, in.mask);
                int lz = -1;
                for (size_t q = 0; q < lz_keys.size(); ++q)
                    if (lz_keys[q] == lkey) lz = static_cast<int>(q);
                if (lz < 0) {
                    lz = static_cast<int>(lz_keys.size());
                    lz_keys.push_back(lkey);
                    if (in.b >= 0) {
                        snprintf(a3, sizeof a3, "%%rx%d", in.b);
                        snprintf(a4, sizeof a4, "%%ry%d", in.b);
                    } else {
                        snprintf(a3, sizeof a3, "%%lx%d", in.t2);
                        snprintf(a4, sizeof a4, "%%ly%d", in.t2);
                    }
                    const int id = lz_base + lz;
                    if (in.mask == 0xffffffffu) {
                        body("neg.f64 %%mx%d, %s;", id, a3);
                        if (cx) body("mov.f64 %%my%d, %s;", id, a4);
                    } else {
                        maskpred("%p1", in.mask);
                        body("neg.f64 %%fa, %s;", a3);
                        body("selp.f64 %%mx%d, %%fa, 0d0000000000000000, %%p1;", id);
                        if (cx) body("selp.f64 %%my%d, %s, 0d0000000000000000, %%p1;", id, a4);
                    }
                }
                {
                    const int id = lz_base + lz;  // mx = -l.x (masked), my = +l.y (masked)
                    if (cx) {
                        body("fma.rn.f64 %s, %%mx%d, %%ux%d, %s;", RX(in.a), id, in.t, RX(in.a));
                        body("fma.rn.f64 %s, %%my%d, %%uy%d, %s;", RX(in.a), id, in.t, RX(in.a));
                        body("fma.rn.f64 %s, %%mx%d, %%uy%d, %s;", RY(in.a), id, in.t, RY(in.a));
                        body("neg.f64 %%fb, %%my%d;", id);
                        body("fma.rn.f64 %s, %%fb, %%ux%d, %s;", RY(in.a), in.t, RY(in.a));
                    } else {
                        body("fma.rn.f64 %s, %%mx%d, %%ux%d, %s;", RX(in.a), id, in.t, RX(in.a));
                    }
                }
                break;
            }
            case OP_DUMP: {
                maskpred("%p1", in.mask);
                if (cx)
                    body("@%%p1 st.shared.v2.f64 [%%sl+%d], {%s, %s};", in.t * kLanes * VS, RX(in.a), RY(in.a));
                else
                    body("@%%p1 st.shared.f64 [%%sl+%d], %s;", in.t * kLanes * VS, RX(in.a));
                break;
            }
            case OP_DUMPT: {
                maskpred("%p1", in.mask);
                tabload("%t1", in.tab);
                body("mad.lo.u32 %%t2, %%t1, %d, %%su;", VS);
                if (cx)
                    body("@%%p1 st.shared.v2.f64 [%%t2], {%s, %s};", RX(in.a), RY(in.a));
                else
                    body("@%%p1 st.shared.f64 [%%t2], %s;", RX(in.a));
                break;
            }
            case OP_XMUL:
            case OP_XMUL_S: {
                if (in.op == OP_XMUL) {
                    snprintf(a3, sizeof a3, "%%rx%d", in.b);
                    snprintf(a4, sizeof a4, "%%ry%d", in.b);
                } else {
                    maskpred("%p1", in.mask);
                    body("mov.f64 %%fc, 0d0000000000000000;");
                    if (cx) {
                        body("mov.f64 %%fm, 0d0000000000000000;");
                        body("@%%p1 ld.shared.v2.f64 {%%fc, %%fm}, [%%sl+%d];", in.t2 * kLanes * VS);
                    } else {
                        body("@%%p1 ld.shared.f64 %%fc, [%%sl+%d];", in.t2 * kLanes * VS);
                    }
                    snprintf(a3, sizeof a3, "%%fc");
                    snprintf(a4, sizeof a4, "%%fm");
                }
                if (cx) {
                    body("mul.f64 %%xx%d, %s, %s;", in.t, RX(in.a), a3);
                    body("neg.f64 %%fa, %s;", RY(in.a));
                    body("fma.rn.f64 %%xx%d, %%fa, %s, %%xx%d;", in.t, a4, in.t);
                    body("mul.f64 %%xy%d, %s, %s;", in.t, RX(in.a), a4);
                    body("fma.rn.f64 %%xy%d, %s, %s, %%xy%d;", in.t, RY(in.a), a3, in.t);
                } else {
                    body("mul.f64 %%xx%d, %s, %s;", in.t, RX(in.a), a3);
                }
                break;
            }
            case OP_SETX: {
                maskpred("%p1", in.mask);
                body("@%%p1 mov.f64 %s, %%xx%d;", RX(in.a), in.t);
                if (cx) body("@%%p1 mov.f64 %s, %%xy%d;", RY(in.a), in.t);
                break;
            }
            case OP_BCASTX: {
                lz_base += static_cast<int>(lz_keys.size());  // back substitution: multipliers are U entries
                lz_keys.clear();
                snprintf(a1, sizeof a1, "%%ux%d", in.t2);
                snprintf(a3, sizeof a3, "%%xx%d", in.t);
                shfl64(a1, a3, in.b);
                if (cx) {
                    snprintf(a2, sizeof a2, "%%uy%d", in.t2);
                    snprintf(a4, sizeof a4, "%%xy%d", in.t);
                    shfl64(a2, a4, in.b);
                }
                break;
            }
            case OP_MAC_S: {
                maskpred("%p1", in.mask);
                body("mov.f64 %%fc, 0d0000000000000000;");
                if (cx) {
                    body("mov.f64 %%fm, 0d0000000000000000;");
                    body("@%%p1 ld.shared.v2.f64 {%%fc, %%fm}, [%%sl+%d];", in.t2 * kLanes * VS);
                    body("neg.f64 %%fa, %%fc;");
                    body("neg.f64 %%fb, %%fm;");
                    body("fma.rn.f64 %s, %%fa, %%ux%d, %s;", RX(in.a), in.t, RX(in.a));
                    body("fma.rn.f64 %s, %%fm, %%uy%d, %s;", RX(in.a), in.t, RX(in.a));
                    body("fma.rn.f64 %s, %%fa, %%uy%d, %s;", RY(in.a), in.t, RY(in.a));
                    body("fma.rn.f64 %s, %%fb, %%ux%d, %s;", RY(in.a), in.t, RY(in.a));
                } else {
                    body("@%%p1 ld.shared.f64 %%fc, [%%sl+%d];", in.t2 * kLanes * VS);
                    body("neg.f64 %%fa, %%fc;");
                    body("fma.rn.f64 %s, %%fa, %%ux%d, %s;", RX(in.a), in.t, RX(in.a));
                }
                break;
            }
            case OP_STOREX: {
                tabload("%t1", in.tab);
                body("setp.ne.and.u32 %%p1, %%t1, 65535, %%p3;");
                body("mov.f64 %%fa, %s;", RX(in.a));
                if (cx) body("mov.f64 %%fb, %s;", RY(in.a));
                bool any = false;
                for (int l = 0; l < kLanes; ++l) any |= p.tabs[static_cast<size_t>(in.tab2) * kLanes + l] != kNone;
                if (any) {
                    tabload("%t3", in.tab2);
                    body("setp.ne.u32 %%p2, %%t3, 65535;");
                    body("mad.lo.u32 %%t3, %%t3, 8, %%srs;");
                    body("mov.f64 %%fc, 0d3FF0000000000000;");
                    body("@%%p2 ld.shared.f64 %%fc, [%%t3];");
                    body("div.rn.f64 %%fa, %%fa, %%fc;");
                    if (cx) body("div.rn.f64 %%fb, %%fb, %%fc;");
                }
                body("mul.wide.u32 %%rd0, %%t1, %d;", VS);
                body("add.u64 %%rd0, %%rd0, %%xm;");
                if (cx)
                    body("@%%p1 st.global.v2.f64 [%%rd0], {%%fa, %%fb};");
                else
                    body("@%%p1 st.global.f64 [%%rd0], %%fa;");
                break;
            }
            case OP_XBUF: {
                tabload("%t1", in.tab);
                body("setp.ne.u32 %%p1, %%t1, 65535;");
                body("mad.lo.u32 %%t2, %%t1, %d, %%sx;", VS);
                if (cx)
                    body("@%%p1 st.shared.v2.f64 [%%t2], {%s, %s};", RX(in.a), RY(in.a));
                else
                    body("@%%p1 st.shared.f64 [%%t2], %s;", RX(in.a));
                break;
            }
            case OP_SYNC:
                body("bar.warp.sync 0xffffffff;");
                break;
            default:
                break;
        }
    }

    // ---- the kernel around the body ------------------------------------------------------------
    W k;
    k(".version 8.7");
    k(".target sm_100a");
    k(".address_size 64");
    k(".extern .shared .align 16 .b8 rr_smem[];");
    k(".visible .entry %s(", out.name.c_str());
    k("  .param .u64 p_Ax, .param .u64 p_b, .param .u64 p_x, .param .u64 p_tabs, .param .u64 p_dst,");
    k("  .param .u64 p_status, .param .u64 p_growth, .param .u64 p_bad,");
    k("  .param .u64 p_axs, .param .u64 p_bs, .param .u64 p_prof, .param .u32 p_L)");
    k(".maxntid %d, 1, 1", WPC * 32);
    k("{");
    k(".reg .f64 %%rx<%d>;", std::max(1, p.P));
    k(".reg .f64 %%ry<%d>;", std::max(1, p.P));
    k(".reg .f64 %%ux<%d>;", std::max(1, p.n_u));
    k(".reg .f64 %%uy<%d>;", std::max(1, p.n_u));
    k(".reg .f64 %%xx<%d>;", std::max(1, p.n_x));
    k(".reg .f64 %%xy<%d>;", std::max(1, p.n_x));
    k(".reg .f64 %%mx<%d>;", lz_base + static_cast<int>(lz_keys.size()) + 1);
    k(".reg .f64 %%my<%d>;", lz_base + static_cast<int>(lz_keys.size()) + 1);
    k(".reg .f64 %%lx<%d>;", std::max(1, p.n_l));
    k(".reg .f64 %%ly<%d>;", std::max(1, p.n_l));
    k(".reg .f64 %%ix<%d>;", std::max(1, p.n_inv));
    k(".reg .f64 %%iy<%d>;", std::max(1, p.n_inv));
    k(".reg .f64 %%fa, %%fb, %%fc, %%fm, %%fic, %%frs, %%dx, %%dy, %%gmax;");
    k(".reg .b32 %%s<4>;");
    k(".reg .u32 %%t<4>;");
    k(".reg .u32 %%lane, %%warp, %%lanebit, %%sa, %%su, %%sl, %%sb, %%sx, %%srs, %%bad, %%m, %%meff, %%mbase, %%L, %%mstep, %%e;");
    k(".reg .s32 %%d0;");
    k(".reg .u16 %%h0;");
    k(".reg .pred %%p<4>;");
    k(".reg .u64 %%rd<4>;");
    k(".reg .u64 %%Ax, %%b, %%x, %%tabl, %%dst, %%status, %%growth, %%pbad, %%axs, %%bs, %%am, %%bm, %%xm, %%prof;");
    k(".reg .u64 %%ck<5>;");
    k("ld.param.u64 %%Ax, [p_Ax];");
    k("ld.param.u64 %%b, [p_b];");
    k("ld.param.u64 %%x, [p_x];");
    k("ld.param.u64 %%tabl, [p_tabs];");
    k("ld.param.u64 %%dst, [p_dst];");
    k("ld.param.u64 %%status, [p_status];");
    k("ld.param.u64 %%growth, [p_growth];");
    k("ld.param.u64 %%pbad, [p_bad];");
    k("ld.param.u64 %%axs, [p_axs];");
    k("ld.param.u64 %%bs, [p_bs];");
    k("ld.param.u32 %%L, [p_L];");
    k("ld.param.u64 %%prof, [p_prof];");
    k("cvta.to.global.u64 %%Ax, %%Ax;");
    k("cvta.to.global.u64 %%b, %%b;");
    k("cvta.to.global.u64 %%x, %%x;");
    k("cvta.to.global.u64 %%tabl, %%tabl;");
    k("cvta.to.global.u64 %%dst, %%dst;");
    k("cvta.to.global.u64 %%status, %%status;");
    k("cvta.to.global.u64 %%growth, %%growth;");
    k("cvta.to.global.u64 %%pbad, %%pbad;");
    k("mov.u32 %%t0, %%tid.x;");
    k("and.b32 %%lane, %%t0, 31;");
    k("shr.u32 %%warp, %%t0, 5;");
    k("shl.b32 %%lanebit, 1, %%lane;");
    k("mul.wide.u32 %%rd0, %%lane, 2;");
    k("add.u64 %%tabl, %%tabl, %%rd0;");  // this lane's column of every table
    k("mov.u32 %%t0, rr_smem;");
    k("mad.lo.u32 %%sa, %%warp, %zu, %%t0;", out.smem_per_warp);
    k("add.u32 %%su, %%sa, %zu;", off_u);
    k("mad.lo.u32 %%sl, %%lane, %d, %%su;", VS);
    k("add.u32 %%sb, %%sa, %zu;", off_b);
    k("add.u32 %%sx, %%sa, %zu;", off_x);
    k("add.u32 %%srs, %%sa, %zu;", off_rs);
    k("mov.u32 %%t0, %%ctaid.x;");
    k("mad.lo.u32 %%m, %%t0, %d, %%warp;", WPC);
    k("mov.u32 %%t0, %%nctaid.x;");
    k("mul.lo.u32 %%mstep, %%t0, %d;", WPC);
    k("mov.u32 %%t0, %%ctaid.x;");
    k("mul.lo.u32 %%mbase, %%t0, %d;", WPC);  // first system of the CTA in this round (CTA-uniform)
    k("setp.ge.u32 %%p0, %%mbase, %%L;");
    k("@%%p0 bra.uni EXIT;");
    k("SYSTEM:");
    k("setp.lt.u32 %%p3, %%m, %%L;");          // p3: this warp has a system in this round
    k("sub.u32 %%t0, %%L, 1;");
    k("min.u32 %%meff, %%m, %%t0;");           // idle warps recompute the last system (no stores)
    // ---- stage the system: Abuf[dst[e]] = Ax[m][e], bbuf = b[m] ----
    k("mov.u64 %%ck0, %%clock64;");
    k("mov.u64 %%ck2, %%ck0;");
    k("mov.u64 %%ck3, %%ck0;");
    k("mul.wide.u32 %%rd0, %%meff, %d;", VS);
    k("mul.lo.u64 %%am, %%rd0, %%axs;");
    k("add.u64 %%am, %%am, %%Ax;");
    k("mul.lo.u64 %%bm, %%rd0, %%bs;");
    k("add.u64 %%bm, %%bm, %%b;");
    k("mul.wide.u32 %%rd0, %%meff, %d;", VS * n * r);
    k("add.u64 %%xm, %%x, %%rd0;");
    k("mov.u32 %%bad, 0;");
    k("mov.f64 %%gmax, 0d0000000000000000;");
    for (int e0 = 0; e0 < nz; e0 += kLanes) {
        k("add.u32 %%e, %%lane, %d;", e0);
        if (e0 + kLanes > nz) k("setp.lt.u32 %%p1, %%e, %d;", nz);
        else k("setp.eq.u32 %%p1, 0, 0;");
        k("mul.wide.u32 %%rd0, %%e, 4;");
        k("add.u64 %%rd0, %%rd0, %%dst;");
        k("mov.s32 %%d0, -1;");
        k("@%%p1 ld.global.nc.s32 %%d0, [%%rd0];");
        k("setp.ge.and.s32 %%p1, %%d0, 0, %%p1;");
        k("mul.wide.u32 %%rd1, %%e, %d;", VS);
        k("add.u64 %%rd1, %%rd1, %%am;");
        k("mov.u32 %%t1, %%d0;");
        k("mad.lo.u32 %%t2, %%t1, %d, %%sa;", VS);
        if (cx) {
            k("@%%p1 ld.global.nc.v2.f64 {%%fa, %%fb}, [%%rd1];");
            k("@%%p1 st.shared.v2.f64 [%%t2], {%%fa, %%fb};");
        } else {
            k("@%%p1 ld.global.nc.f64 %%fa, [%%rd1];");
            k("@%%p1 st.shared.f64 [%%t2], %%fa;");
        }
    }
    for (int e0 = 0; e0 < n * r; e0 += kLanes) {
        k("add.u32 %%e, %%lane, %d;", e0);
        if (e0 + kLanes > n * r) k("setp.lt.u32 %%p1, %%e, %d;", n * r);
        else k("setp.eq.u32 %%p1, 0, 0;");
        k("mul.wide.u32 %%rd1, %%e, %d;", VS);
        k("add.u64 %%rd1, %%rd1, %%bm;");
        k("mad.lo.u32 %%t2, %%e, %d, %%sb;", VS);
        if (cx) {
            k("@%%p1 ld.global.nc.v2.f64 {%%fa, %%fb}, [%%rd1];");
            k("@%%p1 st.shared.v2.f64 [%%t2], {%%fa, %%fb};");
        } else {
            k("@%%p1 ld.global.nc.f64 %%fa, [%%rd1];");
            k("@%%p1 st.shared.f64 [%%t2], %%fa;");
        }
    }
    k("bar.warp.sync 0xffffffff;");
    k("mov.u64 %%ck1, %%clock64;");
    k.s += body.s;
    // ---- status and growth of this system ----
    for (int o = 16; o >= 1; o >>= 1) {
        k("mov.b64 {%%s0, %%s1}, %%gmax;");
        k("shfl.sync.bfly.b32 %%s2, %%s0, %d, 0x1f, 0xffffffff;", o);
        k("shfl.sync.bfly.b32 %%s3, %%s1, %d, 0x1f, 0xffffffff;", o);
        k("mov.b64 %%fa, {%%s2, %%s3};");
        k("setp.gtu.f64 %%p1, %%fa, %%gmax;");
        k("@%%p1 mov.f64 %%gmax, %%fa;");
    }
    k("ld.global.s32 %%d0, [%%pbad];");
    k("setp.ne.s32 %%p1, %%d0, 0;");
    k("@%%p1 mov.u32 %%bad, 0xfffffffd;");  // KLU_INVALID (-3)
    k("setp.eq.and.u32 %%p1, %%lane, 0, %%p3;");
    k("mul.wide.u32 %%rd0, %%m, 4;");
    k("add.u64 %%rd0, %%rd0, %%status;");
    k("@%%p1 st.global.u32 [%%rd0], %%bad;");
    k("sqrt.rn.f64 %%fa, %%gmax;");
    k("mul.wide.u32 %%rd0, %%m, 8;");
    k("add.u64 %%rd0, %%rd0, %%growth;");
    k("@%%p1 st.global.f64 [%%rd0], %%fa;");
    k("bar.warp.sync 0xffffffff;");
    k("mov.u64 %%ck4, %%clock64;");
    k("setp.ne.u64 %%p2, %%prof, 0;");
    k("setp.eq.and.u32 %%p2, %%lane, 0, %%p2;");
    k("and.pred %%p2, %%p2, %%p3;");
    k("@!%%p2 bra.uni NOPROF;");
    k("cvta.to.global.u64 %%rd1, %%prof;");
    k("sub.u64 %%rd0, %%ck1, %%ck0;");
    k("red.global.add.u64 [%%rd1], %%rd0;");
    k("sub.u64 %%rd0, %%ck2, %%ck1;");
    k("red.global.add.u64 [%%rd1+8], %%rd0;");
    k("sub.u64 %%rd0, %%ck3, %%ck2;");
    k("red.global.add.u64 [%%rd1+16], %%rd0;");
    k("sub.u64 %%rd0, %%ck4, %%ck3;");
    k("red.global.add.u64 [%%rd1+24], %%rd0;");
    k("red.global.add.u64 [%%rd1+32], 1;");
    k("NOPROF:");
    k("add.u32 %%m, %%m, %%mstep;");
    k("add.u32 %%mbase, %%mbase, %%mstep;");
    k("setp.lt.u32 %%p0, %%mbase, %%L;");
    k("@%%p0 bra.uni SYSTEM;");
    k("EXIT:");
    k("ret;");
    k("}");
    out.ptx.swap(k.s);
    return true;
}

}  // namespace rr
}  // namespace kb
