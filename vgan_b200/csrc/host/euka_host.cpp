// euka host side (include/vgan_host.h): damage tables and the warp-striped batch packer.
// No likelihood arithmetic here: the packer only classifies the aligned symbols; the damage tables are the
// reference's own construction (Damage::initDeamProbabilities) evaluated once per process.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "host_internal.h"

namespace {

const unsigned kMinFragment = 15, kMaxFragment = 1000;  // reference src/damage.h:42-43

// position of the substitution a>b in the .prof column order A>C A>G A>T C>A C>G C>T G>A G>C G>T T>A T>C T>G
inline int sub_column(int a, int b) { return a * 3 + (b > a ? b - 1 : b); }

// One side of the profile: per position the 4x4 substitution matrix, diagonal = 1 - sum of the row's rates
// (reference src/damage.cpp:63-93 for the 5' end, :112-137 for the 3' end).  The identity probability is
// accumulated in `double`, the rates stay `long double`, exactly as the reference declares them.
int profile_rows(const char* text, std::vector<std::array<long double, 16>>& rows) {
    rows.clear();
    std::vector<std::array<long double, 12>> rates;
    if (text && text[0]) {
        std::istringstream in(text);
        std::string line;
        if (!std::getline(in, line)) return vgh_fail(VGH_ERR_FORMAT, "empty damage profile");
        while (std::getline(in, line)) {  // reference src/miscfunc.h:107-127
            std::vector<std::string> f;
            std::istringstream ls(line);
            std::string tok;
            while (std::getline(ls, tok, '\t')) f.push_back(tok);
            if (f.size() == 13) f.pop_back();
            if (f.size() != 12) return vgh_fail(VGH_ERR_FORMAT, "line from error profile has %zu fields rather than 12", f.size());
            std::array<long double, 12> r;
            for (int k = 0; k < 12; ++k) r[(size_t)k] = strtold(f[(size_t)k].c_str(), nullptr);
            rates.push_back(r);
        }
        if (rates.empty()) return vgh_fail(VGH_ERR_FORMAT, "damage profile has no rows");
    } else {
        std::array<long double, 12> z;
        z.fill(0.0L);
        rates.push_back(z);  // no profile: every position is the identity (the reference fills 1000 zero rows)
    }
    for (const auto& r : rates) {
        std::array<long double, 16> m;
        m.fill(0.0L);
        for (int a = 0; a < 4; ++a) {
            double ident = 1.0;
            for (int b = 0; b < 4; ++b) {
                if (a == b) continue;
                ident = ident - r[(size_t)sub_column(a, b)];
                m[(size_t)(a * 4 + b)] = r[(size_t)sub_column(a, b)];
            }
            if (ident < 0) return vgh_fail(VGH_ERR_FORMAT, "Error with deamination profile, identity probability is less than 0");
            m[(size_t)(a * 4 + a)] = ident;
        }
        rows.push_back(m);
    }
    if (rows.size() > kMaxFragment) rows.resize(kMaxFragment);
    return VGH_OK;
}

inline int symbol_code(unsigned char c) {
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        case 'N': return VGL_EUKA_SYM_N;
        case '-': return VGL_EUKA_SYM_GAP;
        case 'W': case 'M': case 'K': case 'R': case 'Y': case 'B': case 'D': case 'H': case 'V': return VGL_EUKA_SYM_RARE;
        case 'S': return VGL_EUKA_SYM_SOFTCLIP;
        default: return VGL_EUKA_SYM_OTHER;
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct vgh_euka_tables {
    vgl_euka_tables view{};
    std::vector<double> dist, deam;
    std::vector<int32_t> bin_off, bin_lo, bin_hi;
};

struct vgh_soibean_tables {
    vgl_soibean_tables view{};
    std::vector<uint32_t> bits;
    std::vector<double> deam;
};

struct vgh_euka_packed {
    vgl_euka_batch batch{};
    bool pinned = false;
    void* block = nullptr;
};

extern "C" {

int vgh_euka_tables_build(int32_t n_clades, const double* clade_dist, const int32_t* bin_off, const int32_t* bin_lo,
                          const int32_t* bin_hi, int32_t max_node, const char* deam5p_prof, const char* deam3p_prof,
                          uint32_t minimum_mq, vgh_euka_tables** out) {
    if (!out || n_clades < 1 || !clade_dist || !bin_off || max_node < 0) return vgh_fail(VGH_ERR_INVALID, "bad euka table arguments");
    std::vector<std::array<long double, 16>> p5, p3;
    if (int rc = profile_rows(deam5p_prof, p5)) return rc;
    if (int rc = profile_rows(deam3p_prof, p3)) return rc;
    vgh_euka_tables* T = new vgh_euka_tables();
    const int nb = bin_off[n_clades];
    T->dist.assign(clade_dist, clade_dist + n_clades);
    T->bin_off.assign(bin_off, bin_off + n_clades + 1);
    T->bin_lo.assign(bin_lo, bin_lo + nb);
    T->bin_hi.assign(bin_hi, bin_hi + nb);
    // subDeamDiNuc[L][l]: per original base the 5' row of position l or the 3' row of position L-l-1, whichever has
    // the SMALLER diagonal; ties take the 5' row (combineDeamRates, reference src/damage.cpp:18-36,244)
    const int k5 = (int)p5.size(), k3 = (int)p3.size();
    T->deam.assign((size_t)k5 * k3 * 16, 0.0);
    for (int i5 = 0; i5 < k5; ++i5)
        for (int i3 = 0; i3 < k3; ++i3)
            for (int a = 0; a < 4; ++a) {
                const long double* f1 = &p5[(size_t)i5][(size_t)(a * 4)];
                const long double* f2 = &p3[(size_t)i3][(size_t)(a * 4)];
                const long double mn = f1[a] < f2[a] ? f1[a] : f2[a];
                const long double* f = (f1[a] == mn) ? f1 : f2;
                for (int b = 0; b < 4; ++b) T->deam[((size_t)i5 * k3 + i3) * 16 + a * 4 + b] = (double)f[b];
            }
    vgl_euka_tables& v = T->view;
    v.n_clades = n_clades;
    v.n_bins = nb;
    v.max_node = max_node;
    v.k5 = k5;
    v.k3 = k3;
    v.minimum_mq = minimum_mq;
    v.clade_dist = T->dist.data();
    v.bin_off = T->bin_off.data();
    v.bin_lo = T->bin_lo.data();
    v.bin_hi = T->bin_hi.data();
    v.deam = T->deam.data();
    for (int Q = 0; Q < 100; ++Q) v.qscore[Q] = Q >= 2 ? std::pow(10, ((-1 * Q) * 0.1)) : 0.25;  // Euka::get_qscore_vec
    v.base_freq_log[0] = std::log(0.362815); v.base_freq_log[1] = std::log(0.207743); v.base_freq_log[2] = std::log(0.116809);
    v.base_freq_log[3] = std::log(0.312435); v.base_freq_log[4] = std::log(0.25);  // A C G T N  (src/Euka.cpp:446-450)
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) v.tt_ratio[a][b] = a == b ? 1.0 : 0.02381;
    v.tt_ratio[0][2] = v.tt_ratio[2][0] = v.tt_ratio[1][3] = v.tt_ratio[3][1] = 0.95238;  // transitions (src/Euka.cpp:453-468)
    *out = T;
    return VGH_OK;
}
const vgl_euka_tables* vgh_euka_tables_view(const vgh_euka_tables* t) { return t ? &t->view : nullptr; }

int vgh_soibean_tables_build(int32_t n_rows, int32_t min_id, int32_t n_paths, const uint8_t* member, const char* deam5p_prof,
                             const char* deam3p_prof, int32_t penalty, vgh_soibean_tables** out) {
    if (!out || n_rows < 1 || n_paths < 1 || !member || penalty < 1) return vgh_fail(VGH_ERR_INVALID, "bad soibean table arguments");
    std::vector<std::array<long double, 16>> p5, p3;
    if (int rc = profile_rows(deam5p_prof, p5)) return rc;
    if (int rc = profile_rows(deam3p_prof, p3)) return rc;
    vgh_soibean_tables* T = new vgh_soibean_tables();
    const int W = (n_paths + 31) / 32;
    T->bits.assign((size_t)n_rows * W, 0u);
    for (int n = 0; n < n_rows; ++n)
        for (int p = 0; p < n_paths; ++p)
            if (member[(size_t)n * n_paths + p]) T->bits[(size_t)n * W + (p >> 5)] |= 1u << (p & 31);
    const int k5 = (int)p5.size(), k3 = (int)p3.size();
    T->deam.assign((size_t)k5 * k3 * 16, 0.0);
    for (int i5 = 0; i5 < k5; ++i5)
        for (int i3 = 0; i3 < k3; ++i3)
            for (int a = 0; a < 4; ++a) {
                const long double* f1 = &p5[(size_t)i5][(size_t)(a * 4)];
                const long double* f2 = &p3[(size_t)i3][(size_t)(a * 4)];
                const long double mn = f1[a] < f2[a] ? f1[a] : f2[a];
                const long double* f = (f1[a] == mn) ? f1 : f2;
                for (int b = 0; b < 4; ++b) T->deam[((size_t)i5 * k3 + i3) * 16 + a * 4 + b] = (double)f[b];
            }
    vgl_soibean_tables& v = T->view;
    v.n_rows = n_rows; v.min_id = min_id; v.n_paths = n_paths; v.row_words = W; v.k5 = k5; v.k3 = k3; v.penalty = penalty;
    v.member_bits = T->bits.data();
    v.deam = T->deam.data();
    for (int Q = 0; Q < 100; ++Q) v.qscore[Q] = Q >= 2 ? std::pow(10, ((-1 * Q) * 0.1)) : 0.25;  // Euka::get_qscore_vec (src/soibean.cpp:471)
    *out = T;
    return VGH_OK;
}
const vgl_soibean_tables* vgh_soibean_tables_view(const vgh_soibean_tables* t) { return t ? &t->view : nullptr; }
void vgh_soibean_tables_destroy(vgh_soibean_tables* t) { delete t; }
void vgh_euka_tables_destroy(vgh_euka_tables* t) { delete t; }

}  // extern "C"

// One packer for both tools.  soibean != 0: "mapping" records are edit records (size << 20 | node of the mapping with the
// same index, 0xFFFFF when that mapping does not exist), the fragment-length check applies to the aligned graph string
// (analyse_GAM indexes the damage table with |graph_seq|, reference src/getLCAfromGAM.h:109,327).
static int pack_symbols(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned, int soibean,
                        vgh_euka_packed** out) {
    if (!g || !a || !out || r0 < 0 || r1 > a->n_reads || r0 > r1) return vgh_fail(VGH_ERR_INVALID, "bad arguments");
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    std::vector<int64_t> keep;
    for (int64_t r = r0; r < r1; ++r)
        if (a->identity[r] != 0) keep.push_back(r);  // readGAM3 skips identity == 0 (src/readGAM_Euka.h:73)
    const size_t R = keep.size();
    std::vector<uint32_t> ncols(R), nmaps(R);
    int err = VGH_OK;
    std::string err_msg;
#pragma omp parallel num_threads(n_threads)
    {
        Recon rc;
#pragma omp for schedule(static)
        for (int64_t k = 0; k < (int64_t)R; ++k) {
            const int64_t r = keep[(size_t)k];
            int e = vgh_recon(*g, *a, r, rc);
            const int64_t M = a->map_off[r + 1] - a->map_off[r];
            const int64_t Lseq = soibean ? (int64_t)rc.graph_seq.size() : a->seq_off[r + 1] - a->seq_off[r];
            if (!e) {
                if (M < 1) e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: no mapping", (long long)r);
                else if (rc.graph_seq.size() != rc.read_seq.size())
                    e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: aligned graph / read strings differ in length", (long long)r);
                else if (Lseq < (int64_t)kMinFragment || Lseq > (int64_t)kMaxFragment)
                    e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: fragment of %lld bp is outside [15, 1000] (the reference indexes its damage table out of range)",
                                 (long long)r, (long long)Lseq);
                else if (rc.graph_seq.size() > 4095 || M > 65535)
                    e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: alignment too long", (long long)r);
                else if (a->mapq[r] < 0 || a->mapq[r] > 255)
                    e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: mapping quality %d out of range", (long long)r, a->mapq[r]);
                else {
                    // the damage position walks the read: it must stay inside [0, Lseq) on every aligned base column
                    int64_t nongap = 0;
                    for (char c : rc.read_seq) nongap += c != '-';
                    if (nongap > Lseq) e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: path spells more bases than the read has", (long long)r);
                }
            }
            if (e) {
#pragma omp critical(vgh_euka_pack_err)
                if (!err) { err = e; err_msg = vgh_error_text(); }
                continue;
            }
            if (soibean) {
                for (int64_t i = 0; i < std::min<int64_t>(M, (int64_t)rc.sizes.size()); ++i) {
                    const int64_t node = a->map_node[a->map_off[r] + i];
                    if (node < 0 || node >= (int64_t)VGL_SOIBEAN_NO_NODE) e = vgh_fail(VGH_ERR_BAD_NODE, "read %lld: node id %lld does not fit the edit record", (long long)r, (long long)node);
                }
                if (a->seq_off[r + 1] - a->seq_off[r] < 1) e = vgh_fail(VGH_ERR_BAD_READ, "read %lld: empty sequence", (long long)r);
                if (e) {
#pragma omp critical(vgh_euka_pack_err)
                    if (!err) { err = e; err_msg = vgh_error_text(); }
                    continue;
                }
            }
            ncols[(size_t)k] = (uint32_t)rc.graph_seq.size();
            nmaps[(size_t)k] = soibean ? (uint32_t)rc.sizes.size() : (uint32_t)M;
        }
    }
    if (err) return vgh_fail(err, "%s", err_msg.c_str());

    std::vector<uint32_t> order(R);
    for (size_t k = 0; k < R; ++k) order[k] = (uint32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return ((uint64_t)ncols[x] << 32 | nmaps[x]) < ((uint64_t)ncols[y] << 32 | nmaps[y]);
    });
    const size_t n_groups = (R + VGL_HC_LANES - 1) / VGL_HC_LANES, lanes = n_groups * VGL_HC_LANES;
    std::vector<uint32_t> g_col(n_groups + 1, 0), g_map(n_groups + 1, 0);
    for (size_t gI = 0; gI < n_groups; ++gI) {
        uint32_t mc = 0, mm = 0;
        for (size_t l = 0; l < VGL_HC_LANES && gI * VGL_HC_LANES + l < R; ++l) {
            mc = std::max(mc, ncols[order[gI * VGL_HC_LANES + l]]);
            mm = std::max(mm, nmaps[order[gI * VGL_HC_LANES + l]]);
        }
        g_col[gI + 1] = g_col[gI] + (mc + VGL_EUKA_COL_BLK - 1) / VGL_EUKA_COL_BLK;
        g_map[gI + 1] = g_map[gI] + (mm + VGL_HC_MAP_BLK - 1) / VGL_HC_MAP_BLK;
    }
    const size_t kColBlk = (size_t)VGL_HC_LANES * VGL_EUKA_COL_BLK, kMapBlk = (size_t)VGL_HC_LANES * VGL_HC_MAP_BLK;
    size_t o = 0;
    const size_t o_gc = o; o = align_up(o + (n_groups + 1) * 4, 256);
    const size_t o_gm = o; o = align_up(o + (n_groups + 1) * 4, 256);
    const size_t o_col = o; o = align_up(o + g_col[n_groups] * kColBlk * 2, 256);
    const size_t o_map = o; o = align_up(o + g_map[n_groups] * kMapBlk * 4, 256);
    const size_t o_nc = o; o = align_up(o + lanes * 2, 256);
    const size_t o_nm = o; o = align_up(o + lanes * 2, 256);
    const size_t o_sl = o; o = align_up(o + lanes * 2, 256);
    const size_t o_mq = o; o = align_up(o + lanes, 256);
    const size_t o_fl = o; o = align_up(o + lanes, 256);
    const size_t o_ri = o; o = align_up(o + lanes * 4, 256);
    vgh_euka_packed* P = new vgh_euka_packed();
    P->pinned = pinned != 0;
    const size_t bytes = std::max<size_t>(o, 256);
    if (pinned) {
        if (vgl_host_alloc(&P->block, bytes) != VGL_OK) { delete P; return vgh_fail(VGH_ERR_INVALID, "pinned allocation failed: %s", vgl_last_error()); }
    } else {
        P->block = aligned_alloc(256, bytes);
        if (!P->block) { delete P; return vgh_fail(VGH_ERR_INVALID, "out of host memory"); }
    }
    uint8_t* base = (uint8_t*)P->block;
    memset(base, 0, bytes);
    memcpy(base + o_gc, g_col.data(), (n_groups + 1) * 4);
    memcpy(base + o_gm, g_map.data(), (n_groups + 1) * 4);
    uint16_t* col = (uint16_t*)(base + o_col);
    uint32_t* map = (uint32_t*)(base + o_map);
    uint16_t* n_cols = (uint16_t*)(base + o_nc);
    uint16_t* n_maps = (uint16_t*)(base + o_nm);
    uint16_t* seq_len = (uint16_t*)(base + o_sl);
    uint8_t* mapq = base + o_mq;
    uint8_t* flags = base + o_fl;
    uint32_t* read_index = (uint32_t*)(base + o_ri);
#pragma omp parallel num_threads(n_threads)
    {
        Recon rc;
#pragma omp for schedule(dynamic, 16)
        for (int64_t gI = 0; gI < (int64_t)n_groups; ++gI) {
            uint16_t* gcol = col + (size_t)g_col[(size_t)gI] * kColBlk;
            uint32_t* gmap = map + (size_t)g_map[(size_t)gI] * kMapBlk;
            for (size_t l = 0; l < VGL_HC_LANES; ++l) {
                const size_t lane_id = (size_t)gI * VGL_HC_LANES + l;
                if (lane_id >= R) continue;
                const uint32_t k = order[lane_id];
                const int64_t r = keep[k];
                if (vgh_recon(*g, *a, r, rc)) continue;
                const size_t L = rc.graph_seq.size();
                const int64_t q0 = a->qual_off[r], Lq = a->qual_off[r + 1] - q0;
                for (size_t m = 0; m < L; ++m) {
                    const uint16_t qb = (int64_t)m < Lq ? (uint8_t)a->qual[q0 + (int64_t)m] : 0;  // quality by ALIGNED COLUMN
                    const uint16_t rec = (uint16_t)(qb | (symbol_code((unsigned char)rc.graph_seq[m]) << 8) |
                                                    (symbol_code((unsigned char)rc.read_seq[m]) << 12));
                    gcol[((m / VGL_EUKA_COL_BLK) * VGL_HC_LANES + l) * VGL_EUKA_COL_BLK + m % VGL_EUKA_COL_BLK] = rec;
                }
                const int64_t mm = a->map_off[r];
                const int64_t M = a->map_off[r + 1] - mm;
                for (uint32_t i = 0; i < nmaps[k]; ++i) {
                    uint32_t rec;
                    if (soibean) {
                        // entries past the number of mappings get no support when the two counts differ (getLCAfromGAM.h:156-160)
                        const bool none = (int64_t)rc.sizes.size() != M && (int64_t)i >= M;
                        rec = ((uint32_t)rc.sizes[i] << 20) | (none ? VGL_SOIBEAN_NO_NODE : (uint32_t)a->map_node[mm + i]);
                    } else {
                        rec = (uint32_t)a->map_node[mm + i];
                    }
                    gmap[((size_t)(i / VGL_HC_MAP_BLK) * VGL_HC_LANES + l) * VGL_HC_MAP_BLK + i % VGL_HC_MAP_BLK] = rec;
                }
                n_cols[lane_id] = (uint16_t)L;
                n_maps[lane_id] = (uint16_t)nmaps[k];
                seq_len[lane_id] = (uint16_t)(a->seq_off[r + 1] - a->seq_off[r]);
                mapq[lane_id] = (uint8_t)a->mapq[r];
                flags[lane_id] = a->map_rev[mm] ? 1 : 0;
                read_index[lane_id] = (uint32_t)(r - r0);
            }
        }
    }
    vgl_euka_batch& b = P->batch;
    b.n_reads = (uint32_t)R;
    b.n_groups = (uint32_t)n_groups;
    b.n_col_blks = g_col[n_groups];
    b.n_map_blks = g_map[n_groups];
    b.grp_col_blk = (const uint32_t*)(base + o_gc);
    b.grp_map_blk = (const uint32_t*)(base + o_gm);
    b.col = col; b.map = map; b.n_cols = n_cols; b.n_maps = n_maps; b.seq_len = seq_len;
    b.mapq = mapq; b.flags = flags; b.read_index = read_index;
    *out = P;
    return VGH_OK;
}
extern "C" {

int vgh_euka_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                  vgh_euka_packed** out) {
    return pack_symbols(g, a, r0, r1, n_threads, pinned, 0, out);
}
int vgh_soibean_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                     vgh_euka_packed** out) {
    return pack_symbols(g, a, r0, r1, n_threads, pinned, 1, out);
}

const vgl_euka_batch* vgh_euka_packed_batch(const vgh_euka_packed* p) { return p ? &p->batch : nullptr; }
void vgh_euka_packed_destroy(vgh_euka_packed* p) {
    if (!p) return;
    if (p->block) { if (p->pinned) vgl_host_free(p->block); else free(p->block); }
    delete p;
}

}  // extern "C"
