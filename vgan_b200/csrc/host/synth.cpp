// Synthetic HaploCart-shaped workloads for tests and bench.py.
// The real tables (share/vgan/hcfiles) are FTP downloads that are absent here, so the benchmark
// shapes are generated: a pangenome of backbone nodes and two-allele bubbles whose alternative
// allele is carried by a random subtree of a random haplogroup tree, node sequences of 1..8 bp, and
// reads sampled from random haplogroup paths with substitutions, indels and (optionally) ancient-DNA
// deamination.  Everything is a pure function of the seeds (own xoshiro256**, no <random>
// distributions, so the streams are identical on every platform).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host_internal.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
    double normal() {  // Box-Muller
        const double u1 = std::max(uniform(), 1e-300), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};
inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}
inline char other_base(Rng& rng, char c) {
    char d;
    do { d = kBases[rng.below(4)]; } while (d == c);
    return d;
}

// Empirical quality-byte histogram of the reference fixture test/input_files/alignments/J2a1a1a1.gam
// (8100 bases; raw PHRED bytes 0..60) and its mapping-quality histogram (81 reads).
const int kQualVal[] = {0, 3, 4, 10, 19, 20, 22, 23, 24, 32, 33, 34, 36, 37, 40, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60};
const int kQualCnt[] = {13, 1, 1, 3, 3, 1, 3, 7, 1, 1, 2, 5, 3, 6, 1, 1, 2, 4, 2, 5, 5, 8, 4, 18, 14, 29, 13, 17, 128, 140, 247, 7412};
const int kMapqVal[] = {0, 1, 3, 60};
const int kMapqCnt[] = {5, 11, 40, 25};

template <size_t N>
int draw_hist(Rng& rng, const int (&val)[N], const int (&cnt)[N]) {
    int tot = 0;
    for (size_t i = 0; i < N; ++i) tot += cnt[i];
    int x = (int)rng.below((uint64_t)tot);
    for (size_t i = 0; i < N; ++i) {
        if (x < cnt[i]) return val[i];
        x -= cnt[i];
    }
    return val[N - 1];
}

}  // namespace

struct vgh_synth_hc {
    int32_t n_rows = 0, n_paths = 0, n_nodes = 0;
    vgh_graph graph;
    std::vector<uint8_t> supports;
    std::vector<int64_t> pmap_node;
    std::vector<int32_t> pmap_val;
    std::vector<double> mappab;
    std::string names_txt, parents_txt, children_txt;
    // read-simulation view of the pangenome
    std::vector<int32_t> site_off;   // [n_sites + 1] path coordinate of each site (identical for every path)
    std::vector<int32_t> site_ref;   // node id carried by paths outside the bubble's subtree (or the backbone node)
    std::vector<int32_t> site_alt;   // node id of the alternative allele, -1 for backbone sites
    std::vector<int32_t> site_lo, site_hi;  // preorder range of the paths carrying the alternative allele
    int32_t total_len = 0;
};

extern "C" {

int vgh_synth_hc_create(uint64_t seed, int32_t n_paths, int32_t n_rows, int32_t n_nodes, vgh_synth_hc** out) {
    if (!out || n_paths < 2 || n_nodes < 4 || n_rows < n_nodes + 1) return vgh_fail(VGH_ERR_INVALID, "bad synth arguments");
    Rng rng(seed);
    vgh_synth_hc* S = new vgh_synth_hc();
    S->n_rows = n_rows; S->n_paths = n_paths; S->n_nodes = n_nodes;
    const int P = n_paths;

    // random recursive tree, relabelled in DFS preorder so that every subtree is a contiguous path range
    std::vector<int> parent0(P, -1);
    std::vector<std::vector<int>> kids0(P);
    for (int i = 1; i < P; ++i) {
        // mix of uniform attachment and attachment to a recent node: gives chains ~ 2x deeper than a uniform tree
        const int lo = (rng.uniform() < 0.5) ? std::max(0, i - 8) : 0;
        parent0[i] = lo + (int)rng.below((uint64_t)(i - lo));
        kids0[parent0[i]].push_back(i);
    }
    std::vector<int> pre(P, 0), size(P, 1), parent(P, -1);
    {
        std::vector<int> order;
        order.reserve(P);
        std::vector<int> stack{0};
        while (!stack.empty()) {
            const int v = stack.back();
            stack.pop_back();
            pre[v] = (int)order.size();
            order.push_back(v);
            for (auto it = kids0[v].rbegin(); it != kids0[v].rend(); ++it) stack.push_back(*it);
        }
        std::vector<int> size0(P, 1);
        for (int k = P - 1; k > 0; --k) size0[parent0[order[k]]] += size0[order[k]];
        for (int v = 0; v < P; ++v) {
            size[pre[v]] = size0[v];
            parent[pre[v]] = parent0[v] < 0 ? -1 : pre[parent0[v]];
        }
    }
    char buf[64];
    std::vector<std::string> names(P);
    for (int p = 0; p < P; ++p) { snprintf(buf, sizeof(buf), "hg%05d", p); names[p] = buf; }
    std::vector<std::vector<int>> kids(P);
    for (int p = 1; p < P; ++p) kids[parent[p]].push_back(p);
    for (int p = 0; p < P; ++p) {
        S->names_txt += names[p];
        S->names_txt += '\n';
        // ancestors from the immediate parent up to the root, then the super-root (not a path)
        S->parents_txt += names[p];
        int hop = 0;
        for (int a = parent[p]; a >= 0; a = parent[a], ++hop) {
            S->parents_txt += ' '; S->parents_txt += names[a];
            if (hop == 0 && p % 97 == 5) { S->parents_txt += ' '; S->parents_txt += names[a]; }      // consecutive duplicate
            if (hop == 1 && p % 89 == 7) S->parents_txt += " [note]";                                 // bracket token is dropped
        }
        S->parents_txt += " mt-MRCA\n";
        if (p % 501 == 3) S->parents_txt += "\n";  // blank lines are skipped
        if (!kids[p].empty()) {
            S->children_txt += names[p];
            for (int c : kids[p]) { S->children_txt += ' '; S->children_txt += names[c]; }
            S->children_txt += '\n';
        }
    }
    S->children_txt += "mt-MRCA " + names[0] + "\n";

    // sites: backbone, bubble, backbone, ..., backbone
    const int nb = (n_nodes - 1) / 3;
    const int n_sites = 2 * nb + 1;
    S->site_off.assign((size_t)n_sites + 1, 0);
    S->site_ref.assign((size_t)n_sites, 0);
    S->site_alt.assign((size_t)n_sites, -1);
    S->site_lo.assign((size_t)n_sites, 0);
    S->site_hi.assign((size_t)n_sites, 0);
    std::vector<std::string> node_seq((size_t)n_nodes + 1);
    int next_id = 1;
    auto draw_len = [&rng]() {
        int len = 1;
        while (len < 8 && rng.uniform() < 0.524) ++len;  // 1 + Geom, clipped to 8: mean ~2.1 bp
        return len;
    };
    for (int k = 0; k < n_sites; ++k) {
        const int len = draw_len();
        std::string s((size_t)len, 'A');
        for (char& c : s) c = kBases[rng.below(4)];
        S->site_off[(size_t)k + 1] = S->site_off[(size_t)k] + len;
        S->site_ref[(size_t)k] = next_id;
        node_seq[(size_t)next_id++] = s;
        if (k & 1) {  // bubble: the alternative allele differs in its first base
            std::string alt = s;
            alt[0] = other_base(rng, s[0]);
            S->site_alt[(size_t)k] = next_id;
            node_seq[(size_t)next_id++] = alt;
            const int v = 1 + (int)rng.below((uint64_t)(P - 1));
            S->site_lo[(size_t)k] = v;
            S->site_hi[(size_t)k] = v + size[v];
        }
    }
    S->total_len = S->site_off[(size_t)n_sites];
    while (next_id <= n_nodes) node_seq[(size_t)next_id++] = "N";  // ids never visited by a path

    S->graph.min_id = 1;
    S->graph.n_nodes = n_nodes;
    S->graph.off.assign((size_t)n_nodes + 1, 0);
    for (int id = 1; id <= n_nodes; ++id) {
        S->graph.seq += node_seq[(size_t)id];
        S->graph.off[(size_t)id] = (int64_t)S->graph.seq.size();
    }

    // membership rows (row index == node id); rows 0 and > n_nodes stay empty like the reference's padding rows
    S->supports.assign((size_t)n_rows * (size_t)P, 0);
    for (int k = 0; k < n_sites; ++k) {
        uint8_t* ref_row = S->supports.data() + (size_t)S->site_ref[(size_t)k] * (size_t)P;
        if (S->site_alt[(size_t)k] < 0) {
            memset(ref_row, 1, (size_t)P);
        } else {
            uint8_t* alt_row = S->supports.data() + (size_t)S->site_alt[(size_t)k] * (size_t)P;
            for (int p = 0; p < P; ++p) {
                const bool in = p >= S->site_lo[(size_t)k] && p < S->site_hi[(size_t)k];
                alt_row[p] = in ? 1 : 0;
                ref_row[p] = in ? 0 : 1;
            }
        }
    }
    // parsed_pangenome_mapping: node -> coordinate of its first base; the reference stores coordinate + 1
    for (int k = 0; k < n_sites; ++k) {
        S->pmap_node.push_back(S->site_ref[(size_t)k]);
        S->pmap_val.push_back(S->site_off[(size_t)k] + 1);
        if (S->site_alt[(size_t)k] >= 0) {
            S->pmap_node.push_back(S->site_alt[(size_t)k]);
            S->pmap_val.push_back(S->site_off[(size_t)k] + 1);
        }
    }
    S->mappab.assign((size_t)S->total_len + 2, 1.0);
    for (double& m : S->mappab) {
        const double u = rng.uniform();
        m = u < 0.02 ? 0.25 : (u < 0.05 ? 0.5 : 1.0);
    }
    *out = S;
    return VGH_OK;
}

void vgh_synth_hc_destroy(vgh_synth_hc* s) { delete s; }
const vgh_graph* vgh_synth_hc_graph(const vgh_synth_hc* s) { return &s->graph; }
const uint8_t* vgh_synth_hc_supports(const vgh_synth_hc* s) { return s->supports.data(); }
int64_t vgh_synth_hc_n_pmap(const vgh_synth_hc* s) { return (int64_t)s->pmap_node.size(); }
const int64_t* vgh_synth_hc_pmap_node(const vgh_synth_hc* s) { return s->pmap_node.data(); }
const int32_t* vgh_synth_hc_pmap_val(const vgh_synth_hc* s) { return s->pmap_val.data(); }
int64_t vgh_synth_hc_n_mappab(const vgh_synth_hc* s) { return (int64_t)s->mappab.size(); }
const double* vgh_synth_hc_mappab(const vgh_synth_hc* s) { return s->mappab.data(); }
const char* vgh_synth_hc_names(const vgh_synth_hc* s) { return s->names_txt.c_str(); }
const char* vgh_synth_hc_parents(const vgh_synth_hc* s) { return s->parents_txt.c_str(); }
const char* vgh_synth_hc_children(const vgh_synth_hc* s) { return s->children_txt.c_str(); }
int32_t vgh_synth_hc_total_len(const vgh_synth_hc* s) { return s->total_len; }
int64_t vgh_synth_hc_graph_seq(const vgh_synth_hc* s, const int64_t** seq_off, const char** seq) {
    *seq_off = s->graph.off.data();
    *seq = s->graph.seq.data();
    return s->graph.n_nodes;
}

}  // extern "C"

namespace {

struct Piece { int32_t node; bool rev; int32_t off, len; };

struct ReadParams {
    int32_t fixed_len;    // > 0: graph bases spanned; 0: ancient-DNA fragment lengths (log-normal, mean 60, [30,150])
    double sub_rate, indel_rate;
    int damage;           // C->T at the 5' end / G->A at the 3' end with the reference's dhigh profiles
    double n_rate;        // read bases replaced by 'N'
    double softclip_rate; // reads that start with 1..5 unaligned bases (an insertion edit in front of the first match)
};

inline int draw_length(Rng& rng, const ReadParams& P, int total_len) {
    int L = P.fixed_len;
    if (L <= 0) {
        const double sigma = 0.4, mu = std::log(60.0) - 0.5 * sigma * sigma;
        L = (int)std::lround(std::exp(mu + sigma * rng.normal()));
        L = std::min(150, std::max(30, L));
    }
    return std::min(L, total_len);
}

// Shared read simulator: `walk_fn(rng, r, walk)` lays out the oriented node pieces the read aligns to; the rest
// (substitutions, indels, deamination, edits, qualities) is common.  One RNG stream per read, so the output does
// not depend on the number of threads.
template <typename WalkFn>
vgh_gam* simulate_reads(const vgh_graph& graph, uint64_t seed, int64_t n_reads, const ReadParams& P, WalkFn walk_fn) {
    static const double k5p[5] = {0.329405, 0.221745, 0.187678, 0.161196, 0.144011};  // C>T column of dhigh5p.prof
    static const double k3p[5] = {0.32891, 0.223405, 0.188599, 0.164419, 0.146352};   // G>A column of dhigh3p.prof
#ifdef _OPENMP
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(omp_get_max_threads(), n_reads / 2048 + 1));
#else
    const int n_threads = 1;
#endif
    std::vector<vgh_gam*> parts((size_t)n_threads, nullptr);
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    const int64_t r_begin = n_reads * tid / n_threads, r_end = n_reads * (tid + 1) / n_threads;
    vgh_gam* G = new vgh_gam();
    parts[(size_t)tid] = G;
    std::vector<Piece> walk;
    std::string gseq, rseq, qual;
    std::vector<uint8_t> op;       // per graph base: 0 match, 1 substitution, 2 deleted
    std::vector<uint8_t> ins_len;  // inserted read bases in front of graph base j (0..5)
    std::vector<char> ins_bases;   // [n][5]
    for (int64_t r = r_begin; r < r_end; ++r) {
        Rng rng(seed * 0x9e3779b97f4a7c15ull + (uint64_t)r * 0xd1342543de82ef95ull + 1);
        walk.clear();
        walk_fn(rng, r, walk);
        gseq.clear();
        for (const Piece& w : walk)
            for (int j = 0; j < w.len; ++j) gseq.push_back(graph.base(w.node, w.rev, w.off + j));
        const int n = (int)gseq.size();
        op.assign((size_t)n, 0);
        ins_len.assign((size_t)n, 0);
        ins_bases.resize((size_t)n * 5 + 5);
        std::string bases = gseq;
        for (int j = 0; j < n; ++j) {
            if (rng.uniform() < P.sub_rate) bases[(size_t)j] = other_base(rng, gseq[(size_t)j]);
            if (j > 0 && j < n - 1 && rng.uniform() < P.indel_rate) {
                if (rng.uniform() < 0.5) {
                    const int k = 1 + (int)rng.below(3);
                    ins_len[(size_t)j] = (uint8_t)k;
                    for (int t = 0; t < k; ++t) ins_bases[(size_t)j * 5 + t] = kBases[rng.below(4)];
                } else {
                    op[(size_t)j] = 2;
                }
            }
        }
        if (n > 0 && P.softclip_rate > 0 && rng.uniform() < P.softclip_rate) {
            const int k = 1 + (int)rng.below(5);
            ins_len[0] = (uint8_t)k;
            for (int t = 0; t < k; ++t) ins_bases[(size_t)t] = kBases[rng.below(4)];
        }
        if (P.damage) {
            int rp = 0;  // read position from the 5' end, counting surviving bases only
            int surviving = 0;
            for (int j = 0; j < n; ++j) surviving += op[(size_t)j] != 2;
            for (int j = 0; j < n; ++j) {
                if (op[(size_t)j] == 2) continue;
                const int d5 = rp, d3 = surviving - 1 - rp;
                const double r5 = d5 < 5 ? k5p[d5] : k5p[4] * std::pow(0.5, d5 - 4);
                const double r3 = d3 < 5 ? k3p[d3] : k3p[4] * std::pow(0.5, d3 - 4);
                if (bases[(size_t)j] == 'C' && rng.uniform() < r5) bases[(size_t)j] = 'T';
                else if (bases[(size_t)j] == 'G' && rng.uniform() < r3) bases[(size_t)j] = 'A';
                ++rp;
            }
        }
        if (P.n_rate > 0)
            for (int j = 0; j < n; ++j)
                if (op[(size_t)j] != 2 && rng.uniform() < P.n_rate) bases[(size_t)j] = 'N';
        for (int j = 0; j < n; ++j)
            if (op[(size_t)j] != 2 && bases[(size_t)j] != gseq[(size_t)j]) op[(size_t)j] = 1;
        // emit mappings / edits
        rseq.clear();
        int j = 0;
        for (const Piece& w : walk) {
            G->map_node.push_back(w.node);
            G->map_offset.push_back(w.off);
            G->map_rev.push_back(w.rev ? 1 : 0);
            const int end = j + w.len;
            while (j < end) {
                if (ins_len[(size_t)j]) {
                    const int k = ins_len[(size_t)j];
                    G->edit_from.push_back(0);
                    G->edit_to.push_back(k);
                    G->edit_seq.append(&ins_bases[(size_t)j * 5], (size_t)k);
                    G->edit_seq_off.push_back((int64_t)G->edit_seq.size());
                    rseq.append(&ins_bases[(size_t)j * 5], (size_t)k);
                    ins_len[(size_t)j] = 0;
                    continue;
                }
                const uint8_t kind = op[(size_t)j];
                int e = j + 1;
                while (e < end && op[(size_t)e] == kind && !ins_len[(size_t)e]) ++e;
                const int len = e - j;
                if (kind == 2) {
                    G->edit_from.push_back(len);
                    G->edit_to.push_back(0);
                } else {
                    G->edit_from.push_back(len);
                    G->edit_to.push_back(len);
                    if (kind == 1) G->edit_seq.append(bases, (size_t)j, (size_t)len);
                    rseq.append(bases, (size_t)j, (size_t)len);
                }
                G->edit_seq_off.push_back((int64_t)G->edit_seq.size());
                j = e;
            }
            G->edit_off.push_back((int64_t)G->edit_from.size());
        }
        qual.resize(rseq.size());
        for (char& q : qual) q = (char)draw_hist(rng, kQualVal, kQualCnt);
        G->seq += rseq; G->seq_off.push_back((int64_t)G->seq.size());
        G->qual += qual; G->qual_off.push_back((int64_t)G->qual.size());
        G->mapq.push_back(draw_hist(rng, kMapqVal, kMapqCnt));
        G->identity.push_back(rng.uniform() < 0.001 ? 0.0 : 1.0);  // a few "unmapped" reads the hot loops must skip
        G->paired.push_back(0);
        G->map_off.push_back((int64_t)G->map_node.size());
    }
    }  // omp parallel
    // concatenate the per-thread pieces in read order
    vgh_gam* G = parts[0];
    for (int t = 1; t < n_threads; ++t) {
        vgh_gam* Q = parts[(size_t)t];
        const int64_t s0 = (int64_t)G->seq.size(), q0 = (int64_t)G->qual.size(), m0 = (int64_t)G->map_node.size();
        const int64_t e0 = (int64_t)G->edit_from.size(), es0 = (int64_t)G->edit_seq.size();
        G->seq += Q->seq; G->qual += Q->qual; G->edit_seq += Q->edit_seq;
        for (size_t k = 1; k < Q->seq_off.size(); ++k) G->seq_off.push_back(Q->seq_off[k] + s0);
        for (size_t k = 1; k < Q->qual_off.size(); ++k) G->qual_off.push_back(Q->qual_off[k] + q0);
        for (size_t k = 1; k < Q->map_off.size(); ++k) G->map_off.push_back(Q->map_off[k] + m0);
        for (size_t k = 1; k < Q->edit_off.size(); ++k) G->edit_off.push_back(Q->edit_off[k] + e0);
        for (size_t k = 1; k < Q->edit_seq_off.size(); ++k) G->edit_seq_off.push_back(Q->edit_seq_off[k] + es0);
        G->mapq.insert(G->mapq.end(), Q->mapq.begin(), Q->mapq.end());
        G->identity.insert(G->identity.end(), Q->identity.begin(), Q->identity.end());
        G->paired.insert(G->paired.end(), Q->paired.begin(), Q->paired.end());
        G->map_node.insert(G->map_node.end(), Q->map_node.begin(), Q->map_node.end());
        G->map_offset.insert(G->map_offset.end(), Q->map_offset.begin(), Q->map_offset.end());
        G->map_rev.insert(G->map_rev.end(), Q->map_rev.begin(), Q->map_rev.end());
        G->edit_from.insert(G->edit_from.end(), Q->edit_from.begin(), Q->edit_from.end());
        G->edit_to.insert(G->edit_to.end(), Q->edit_to.begin(), Q->edit_to.end());
        delete Q;
    }
    G->finish();
    return G;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// euka-shaped workload: clades own consecutive node ranges (one linear mitogenome chopped
// into nodes of up to 32 bp), 10-15 overlapping bins per clade, a pairwise distance per clade.
// ------------------------------------------------------------------------------------------------
struct vgh_synth_euka {
    int32_t n_clades = 0;
    vgh_graph graph;
    std::vector<double> dist;
    std::vector<int32_t> snode, enode;      // node range of each clade
    std::vector<int32_t> bin_off, bin_lo, bin_hi;
    std::vector<int32_t> node_off;          // coordinate of each node inside its clade's genome
    std::vector<int32_t> clade_len;         // genome length of each clade
    std::vector<double> clade_cdf;          // empty: clades equally likely; else cumulative weights (config 4: node-range size)
};

extern "C" {

// Reads: `fixed_len` > 0 gives reads spanning exactly that many graph bases (config 2: 100); 0 gives the
// ancient-DNA fragment-length distribution of config 3 (log-normal, mean 60, clipped to [30,150]).
// `damage` != 0 applies C->T at the 5' end and G->A at the 3' end with the per-position rates of the
// reference's share/vgan/damageProfiles/dhigh{5p,3p}.prof (first five positions; the rate halves with
// every further position).  `origin_path` >= 0 samples every read from that path (a "sample" of one
// haplogroup); -1 picks a random path per read.
int vgh_synth_hc_reads(const vgh_synth_hc* S, uint64_t seed, int64_t n_reads, int32_t fixed_len, double sub_rate,
                       double indel_rate, int damage, int32_t origin_path, vgh_gam** out) {
    if (!S || !out || n_reads < 0) return vgh_fail(VGH_ERR_INVALID, "bad synth arguments");
    const ReadParams P{fixed_len, sub_rate, indel_rate, damage, 0.0, 0.0};
    const int n_sites = (int)S->site_ref.size();
    *out = simulate_reads(S->graph, seed, n_reads, P, [&](Rng& rng, int64_t, std::vector<Piece>& walk) {
        const int p = origin_path >= 0 ? origin_path % S->n_paths : (int)rng.below((uint64_t)S->n_paths);
        const int L = draw_length(rng, P, S->total_len);
        const int s = (int)rng.below((uint64_t)(S->total_len - L + 1));
        const bool rev = rng.uniform() < 0.5;
        const int k0 = (int)(std::upper_bound(S->site_off.begin(), S->site_off.end(), s) - S->site_off.begin()) - 1;
        for (int k = k0; k < n_sites && S->site_off[(size_t)k] < s + L; ++k) {
            const int a = std::max(s, S->site_off[(size_t)k]), b = std::min(s + L, S->site_off[(size_t)k + 1]);
            const bool alt = S->site_alt[(size_t)k] >= 0 && p >= S->site_lo[(size_t)k] && p < S->site_hi[(size_t)k];
            const int node = alt ? S->site_alt[(size_t)k] : S->site_ref[(size_t)k];
            const int nlen = S->site_off[(size_t)k + 1] - S->site_off[(size_t)k];
            const int fo = a - S->site_off[(size_t)k];
            walk.push_back(Piece{node, rev, rev ? nlen - (fo + (b - a)) : fo, b - a});
        }
        if (rev) std::reverse(walk.begin(), walk.end());
    });
    return VGH_OK;
}

int vgh_synth_euka_create(uint64_t seed, int32_t n_clades, int32_t genome_len, vgh_synth_euka** out) {
    if (!out || n_clades < 1 || genome_len < 400) return vgh_fail(VGH_ERR_INVALID, "bad synth arguments");
    Rng rng(seed);
    vgh_synth_euka* S = new vgh_synth_euka();
    S->n_clades = n_clades;
    S->graph.min_id = 1;
    S->graph.off.push_back(0);
    S->bin_off.push_back(0);
    S->node_off.push_back(0);  // node id 0 does not exist; keeps node_off indexable by id
    int next_id = 1;
    for (int c = 0; c < n_clades; ++c) {
        S->dist.push_back(0.05 + 0.15 * rng.uniform());
        const int glen = genome_len + (int)rng.below((uint64_t)(genome_len / 8 + 1));
        S->snode.push_back(next_id);
        int pos = 0;
        while (pos < glen) {
            const int len = std::min(glen - pos, 1 + (int)rng.below(32));
            for (int j = 0; j < len; ++j) S->graph.seq.push_back(kBases[rng.below(4)]);
            S->graph.off.push_back((int64_t)S->graph.seq.size());
            S->node_off.push_back(pos);
            pos += len;
            ++next_id;
        }
        S->enode.push_back(next_id - 1);
        S->clade_len.push_back(glen);
        // 10..15 bins that overlap their neighbours by ~10 % of a bin (as in share/vgan/euka_dir/euka_db.bins)
        const int nb = 10 + (int)rng.below(6);
        const int first = S->snode[(size_t)c], last = S->enode[(size_t)c], span = last - first + 1;
        for (int j = 0; j < nb; ++j) {
            const int lo = first + (int)((int64_t)span * j / nb);
            const int hi = std::min(last, first + (int)((int64_t)span * (j + 1) / nb) + span / (10 * nb));
            S->bin_lo.push_back(std::max(first, lo - (j ? span / (10 * nb) : 0)));
            S->bin_hi.push_back(hi);
        }
        S->bin_off.push_back((int32_t)S->bin_lo.size());
    }
    S->graph.n_nodes = next_id - 1;
    *out = S;
    return VGH_OK;
}
// the clade table (dist, snode..enode) and the bins are taken verbatim from the reference's
// share/vgan/euka_dir/euka_db.clade / euka_db.bins; only the node sequences are synthetic (1-4 random bases per node, a
// clade's genome is the chain of its nodes; ids outside every clade get one base).  A read's clade is drawn in
// proportion to the clade's node-range size.
int vgh_synth_euka_create_from_tables(uint64_t seed, int32_t n_clades, const double* dist, const int32_t* snode, const int32_t* enode,
                                      const int32_t* bin_off, const int32_t* bin_lo, const int32_t* bin_hi, vgh_synth_euka** out) {
    if (!out || n_clades < 1 || !dist || !snode || !enode || !bin_off || !bin_lo || !bin_hi) return vgh_fail(VGH_ERR_INVALID, "bad synth arguments");
    int32_t max_id = 0;
    for (int c = 0; c < n_clades; ++c) {
        if (snode[c] < 1 || enode[c] < snode[c]) return vgh_fail(VGH_ERR_INVALID, "clade %d: bad node range", c);
        max_id = std::max(max_id, enode[c]);
    }
    Rng rng(seed);
    vgh_synth_euka* S = new vgh_synth_euka();
    S->n_clades = n_clades;
    S->graph.min_id = 1;
    S->graph.n_nodes = max_id;
    std::vector<uint8_t> len((size_t)max_id + 1, 1);
    std::vector<int32_t> owner((size_t)max_id + 1, -1);
    for (int c = 0; c < n_clades; ++c)
        for (int32_t id = snode[c]; id <= enode[c]; ++id) owner[(size_t)id] = c;   // (ranges do not overlap in euka_db.clade)
    S->graph.off.reserve((size_t)max_id + 1);
    S->graph.off.push_back(0);
    S->node_off.assign((size_t)max_id + 1, 0);
    for (int32_t id = 1; id <= max_id; ++id) {
        const int l = owner[(size_t)id] >= 0 ? 1 + (int)rng.below(4) : 1;
        for (int j = 0; j < l; ++j) S->graph.seq.push_back(kBases[rng.below(4)]);
        S->graph.off.push_back((int64_t)S->graph.seq.size());
        len[(size_t)id] = (uint8_t)l;
    }
    double tot = 0;
    for (int c = 0; c < n_clades; ++c) {
        S->dist.push_back(dist[c]);
        S->snode.push_back(snode[c]);
        S->enode.push_back(enode[c]);
        int pos = 0;
        for (int32_t id = snode[c]; id <= enode[c]; ++id) { S->node_off[(size_t)id] = pos; pos += len[(size_t)id]; }
        S->clade_len.push_back(pos);
        tot += (double)(enode[c] - snode[c] + 1);
        S->clade_cdf.push_back(tot);
    }
    for (auto& w : S->clade_cdf) w /= tot;
    S->bin_off.assign(bin_off, bin_off + n_clades + 1);
    S->bin_lo.assign(bin_lo, bin_lo + bin_off[n_clades]);
    S->bin_hi.assign(bin_hi, bin_hi + bin_off[n_clades]);
    *out = S;
    return VGH_OK;
}
void vgh_synth_euka_destroy(vgh_synth_euka* s) { delete s; }
const vgh_graph* vgh_synth_euka_graph(const vgh_synth_euka* s) { return &s->graph; }
int64_t vgh_synth_euka_graph_seq(const vgh_synth_euka* s, const int64_t** seq_off, const char** seq) {
    *seq_off = s->graph.off.data();
    *seq = s->graph.seq.data();
    return s->graph.n_nodes;
}
int32_t vgh_synth_euka_n_clades(const vgh_synth_euka* s) { return s->n_clades; }
const double* vgh_synth_euka_dist(const vgh_synth_euka* s) { return s->dist.data(); }
const int32_t* vgh_synth_euka_bin_off(const vgh_synth_euka* s) { return s->bin_off.data(); }
const int32_t* vgh_synth_euka_bin_lo(const vgh_synth_euka* s) { return s->bin_lo.data(); }
const int32_t* vgh_synth_euka_bin_hi(const vgh_synth_euka* s) { return s->bin_hi.data(); }

// Reads of a random clade; substitutions at a quarter of the clade's pairwise distance, plus the knobs above.
int vgh_synth_euka_reads(const vgh_synth_euka* S, uint64_t seed, int64_t n_reads, int32_t fixed_len, double indel_rate,
                         int damage, double n_rate, double softclip_rate, vgh_gam** out) {
    if (!S || !out || n_reads < 0) return vgh_fail(VGH_ERR_INVALID, "bad synth arguments");
    ReadParams P{fixed_len, 0.03, indel_rate, damage, n_rate, softclip_rate};
    *out = simulate_reads(S->graph, seed, n_reads, P, [&](Rng& rng, int64_t, std::vector<Piece>& walk) {
        int c = (int)rng.below((uint64_t)S->n_clades);
        if (!S->clade_cdf.empty()) {
            c = (int)(std::lower_bound(S->clade_cdf.begin(), S->clade_cdf.end(), rng.uniform()) - S->clade_cdf.begin());
            if (c >= S->n_clades) c = S->n_clades - 1;
        }
        const int glen = S->clade_len[(size_t)c];
        const int L = draw_length(rng, P, glen);
        const int s = (int)rng.below((uint64_t)(glen - L + 1));
        const bool rev = rng.uniform() < 0.5;
        // first node of the clade whose coordinate range reaches s
        int lo = S->snode[(size_t)c], hi = S->enode[(size_t)c];
        while (lo < hi) {
            const int mid = (lo + hi + 1) / 2;
            if (S->node_off[(size_t)mid] <= s) lo = mid; else hi = mid - 1;
        }
        for (int id = lo; id <= S->enode[(size_t)c] && S->node_off[(size_t)id] < s + L; ++id) {
            const int nlen = (int)S->graph.length(id);
            const int a = std::max(s, S->node_off[(size_t)id]), b = std::min(s + L, S->node_off[(size_t)id] + nlen);
            const int fo = a - S->node_off[(size_t)id];
            walk.push_back(Piece{id, rev, rev ? nlen - (fo + (b - a)) : fo, b - a});
        }
        if (rev) std::reverse(walk.begin(), walk.end());
    });
    return VGH_OK;
}

}  // extern "C"
