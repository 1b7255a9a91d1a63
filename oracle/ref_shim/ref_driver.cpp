// =====================================================================================
// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE'S OWN hot-path code.
//
// oracle/_ref/libvgan_ref.so is built (oracle/ref_shim/Makefile) from the reference's sources
// where they lie under /root/reference/src -- update_likelihood.cpp (+ process_mapping.cpp),
// get_p_obs_base.cpp, get_posterior.cpp, haplocart_functions.cpp, vgan_utils.h, miscfunc.h,
// NodeInfo.cpp, AlignmentInfo.cpp, rmdup.cpp, damage.cpp -- compiled UNCHANGED against the header
// shim in oracle/ref_shim/include (stand-ins for libvg / protobuf / bdsg / libgab declarations).
// Nothing of the reference is copied into this repository.
//
// This file only marshals flat arrays (the same orc_* structs the oracle takes) into the
// reference's types (vg::Alignment, AlignmentInfo, NodeInfo, bdsg::ODGI, std::map tables) and
// calls the reference's functions.  The loops that are not functions in the reference (the OpenMP
// read loop of Haplocart::run, src/HaploCart.cpp:408-424) are restated here and say so.
//
// Used by tests/ (oracle == reference), tests/golden/make_ref_golden.py and bench.py's
// --impl reference / cpu_baseline legs.  Never by the product.
// =====================================================================================
#include "HaploCart.h"
#include "miscfunc.h"
#include "vgan_utils.h"
#include "Dup_Remover.h"
#include "damage.h"
#include "Euka.h"          // pulls in readGAM_Euka.h (the readGAM3 lambda) and MCMC.h (computeBaseLogLike)
#include "soibean.h"       // pulls in getLCAfromGAM.h (analyse_GAM)
#include "vgan_haplocart.hpp"   // the PRODUCT's C++ facade: its FlatAlignments::assign is compiled here against the reference's AlignmentInfo
#ifdef _OPENMP
#include <omp.h>
#endif
#include <sys/wait.h>
#include <unistd.h>

// ---- pieces of the reference that live in files this build cannot compile (CLI / vg subcommands) ----
Haplocart::Haplocart() {}   // src/HaploCart.cpp:24 (empty there too)
Haplocart::~Haplocart() {}
int main_filter(int, char**) { std::abort(); }
int main_gamsort(int, char**) { std::abort(); }
int main_giraffe(int, char**) { std::abort(); }
// Euka.cpp / soibean.cpp hold the CLI and the graph loaders (gbwt, vg subcommands) and cannot be compiled here; the
// three small members of theirs that the path needs are restated: the empty constructors (src/Euka.cpp:19-25) and
// Euka::get_qscore_vec (src/Euka.cpp:38-51: 10^(-Q/10) for Q >= 2, else 0.25 -- quirk Q7, qs[2] = 0.631).
Euka::Euka() {}
Euka::~Euka() {}
vector<double> Euka::get_qscore_vec() {
    vector<double> qscore_vec;
    for (int Q = 0; Q < 100; ++Q) {
        if (Q >= 2) qscore_vec.emplace_back(pow(10, ((-1 * Q) * 0.1)));
        else qscore_vec.emplace_back(0.25);
    }
    return qscore_vec;
}

extern "C" {

typedef struct { int64_t n_nodes; int64_t min_id; const int64_t* seq_off; const char* seq; } orc_graph;
typedef struct {
    int64_t n_reads;
    const int64_t* seq_off; const char* seq;
    const int64_t* qual_off; const char* qual;
    const int32_t* mapq; const double* identity;
    const int64_t* map_off; const int64_t* map_node; const int64_t* map_offset; const uint8_t* map_rev;
    const int64_t* edit_off; const int32_t* edit_from; const int32_t* edit_to;
    const int64_t* edit_seq_off; const char* edit_seq;
} orc_reads;
typedef struct {
    int32_t n_rows; int32_t n_paths; const uint8_t* supports;
    int64_t n_pmap; const int64_t* pmap_node; const int32_t* pmap_val;
    int64_t n_mappab; const double* mappab;
    double background_error_prob; int32_t is_consensus_fasta; int32_t min_id;
} orc_hc_tables;

}  // extern "C"

namespace {

bdsg::ODGI make_graph(const orc_graph* g) {
    std::vector<std::string> seqs((size_t)g->n_nodes);
    for (int64_t i = 0; i < g->n_nodes; ++i) seqs[(size_t)i].assign(g->seq + g->seq_off[i], g->seq + g->seq_off[i + 1]);
    bdsg::ODGI og;
    og.shim_set(g->min_id, std::move(seqs));
    return og;
}

vg::Alignment make_alignment(const orc_reads* R, int64_t r, int pad_quality = 0) {
    vg::Alignment a;
    a.set_sequence(std::string(R->seq + R->seq_off[r], R->seq + R->seq_off[r + 1]));
    // pad_quality zero bytes make reads past the end of the quality string defined (euka and soibean index the
    // quality by ALIGNED column, which runs past the string for reads with deletions: src/readGAM_Euka.h:264,286)
    a.set_quality(std::string(R->qual + R->qual_off[r], R->qual + R->qual_off[r + 1]) + std::string((size_t)pad_quality, '\0'));
    a.set_mapping_quality(R->mapq[r]);
    a.set_identity(R->identity[r]);
    vg::Path* p = a.mutable_path();
    for (int64_t m = R->map_off[r]; m < R->map_off[r + 1]; ++m) {
        vg::Mapping* mp = p->add_mapping();
        mp->mutable_position()->set_node_id(R->map_node[m]);
        mp->mutable_position()->set_offset(R->map_offset[m]);
        mp->mutable_position()->set_is_reverse(R->map_rev[m] != 0);
        mp->set_rank(m - R->map_off[r] + 1);
        for (int64_t e = R->edit_off[m]; e < R->edit_off[m + 1]; ++e) {
            vg::Edit* ed = mp->add_edit();
            ed->set_from_length(R->edit_from[e]);
            ed->set_to_length(R->edit_to[e]);
            ed->set_sequence(std::string(R->edit_seq + R->edit_seq_off[e], R->edit_seq + R->edit_seq_off[e + 1]));
        }
    }
    return a;
}

// The assignments of the readGAM lambda (src/readGAM.h:41-46).
AlignmentInfo* make_info(const vg::Alignment& a, bool paired = false) {
    AlignmentInfo* ai = new AlignmentInfo();
    ai->seq = a.sequence();
    ai->path = a.path();
    ai->mapping_quality = a.mapping_quality();
    ai->quality_scores = a.quality();
    ai->is_paired = paired;
    ai->identity = a.identity();
    return ai;
}

struct HcState {
    std::map<const std::string, int> pangenome_map;
    std::vector<NodeInfo*> nodevector;
    std::vector<double> mappabilities, qscore_vec, incorrect_mapping_vec;
    int nbpaths = 0, minid = 1;
    double bep = 1e-4;
    bool is_fasta = false;
    ~HcState() { for (auto* n : nodevector) delete n; }
};

// Tables as Haplocart::run holds them (src/HaploCart.cpp:279-286): nodevector[k] is node minid + k and
// carries row (minid + k) of path_supports (src/readPathHandleGraph.cpp:21-34); pangenome_map is keyed
// by the decimal node id (src/load.cpp:27-41); the two 100-entry vectors come from the reference's own
// get_qscore_vec / precompute_incorrect_mapping_probs.
void build_state(const orc_hc_tables* T, HcState& S) {
    Haplocart hc;
    S.nbpaths = T->n_paths;
    S.minid = T->min_id;
    S.bep = T->background_error_prob;
    S.is_fasta = T->is_consensus_fasta != 0;
    for (int64_t i = 0; i < T->n_pmap; ++i) S.pangenome_map.insert(std::make_pair(std::to_string(T->pmap_node[i]), (int)T->pmap_val[i]));
    S.mappabilities.assign(T->mappab, T->mappab + T->n_mappab);
    S.qscore_vec = get_qscore_vec();
    S.incorrect_mapping_vec = hc.precompute_incorrect_mapping_probs();
    for (int64_t id = T->min_id; id < T->n_rows; ++id) {
        NodeInfo* n = new NodeInfo(id, T->n_paths, 0);
        for (int j = 0; j < T->n_paths; ++j) n->pathsgo[j] = T->supports[(size_t)id * (size_t)T->n_paths + (size_t)j] != 0;
        S.nodevector.emplace_back(n);
    }
}

}  // namespace

extern "C" {

int ref_sizeof_long_double(void) { return (int)sizeof(long double); }

// reconstruct_graph_sequence (src/vgan_utils.h:6-79), the reference's own.
int ref_reconstruct(const orc_graph* g, const orc_reads* R, int64_t r, char* graph_seq, char* read_seq, int32_t* sizes,
                    int32_t cap, int32_t* n_graph, int32_t* n_read, int32_t* n_sizes) {
    try {
        const bdsg::ODGI graph = make_graph(g);
        const vg::Alignment a = make_alignment(R, r);
        const auto t = reconstruct_graph_sequence(graph, a.path(), a.sequence());
        const std::string& G = std::get<0>(t);
        const std::string& P = std::get<1>(t);
        const std::vector<int>& S = std::get<2>(t);
        if ((int64_t)G.size() > cap || (int64_t)P.size() > cap || (int64_t)S.size() > cap) return -5;
        memcpy(graph_seq, G.data(), G.size());
        memcpy(read_seq, P.data(), P.size());
        for (size_t i = 0; i < S.size(); ++i) sizes[i] = S[i];
        *n_graph = (int32_t)G.size(); *n_read = (int32_t)P.size(); *n_sizes = (int32_t)S.size();
        return 0;
    } catch (...) { return -1; }
}

// One read through Haplocart::update (src/update_likelihood.cpp:5-16) starting from a zero vector.
// NOTE the reference's quality loop reads past the end of the quality string for every mapping after the
// first (src/update_likelihood.cpp:40-44) -- undefined behaviour.  With pad_quality != 0 the
// driver makes that read DEFINED for this run by appending zero bytes to the AlignmentInfo's quality string
// (the value the oracle and the product define for those bytes); with 0 the string is left as it is and the
// result depends on the heap.
int ref_hc_read(const orc_hc_tables* T, const orc_graph* g, const orc_reads* R, int64_t r, int pad_quality, long double* ll_out) {
    HcState S;
    build_state(T, S);
    const bdsg::ODGI graph = make_graph(g);
    Haplocart hc;
    const vg::Alignment a = make_alignment(R, r);
    AlignmentInfo* ai = make_info(a);
    if (pad_quality) ai->quality_scores.append((size_t)pad_quality, '\0');
    const std::vector<long double> empty_vec((size_t)S.nbpaths, 0);
    const std::vector<long double> ll = hc.update(0, S.pangenome_map, S.nodevector, ai, empty_vec, S.qscore_vec, S.mappabilities, S.nbpaths,
                                                   true, S.is_fasta, S.bep, S.incorrect_mapping_vec, 1, S.minid, S.is_fasta, 1, graph);
    for (int j = 0; j < S.nbpaths; ++j) ll_out[j] = ll[(size_t)j];
    delete ai;
    return 0;
}

// The read loop of Haplocart::run (src/HaploCart.cpp:408-424), restated here because it is not a function in
// the reference: skip identity < 1e-10, update() from the zero vector, add into final_vec (critical section),
// then std::max_element.  `use_background_error_prob` and `is_consensus_fasta` are both set for FASTA input
// (src/HaploCart.cpp:397-401).  per_read_seconds (optional) receives the wall time spent in update().
int ref_hc_run(const orc_hc_tables* T, const orc_graph* g, const orc_reads* R, int64_t r0, int64_t r1, int pad_quality, int n_threads,
               long double* final_out, int32_t* argmax_out, int64_t* used_out, double* loop_seconds) {
    HcState S;
    build_state(T, S);
    const bdsg::ODGI graph = make_graph(g);
    Haplocart hc;
    std::vector<AlignmentInfo*> algnvector;
    for (int64_t r = r0; r < r1; ++r) {
        AlignmentInfo* ai = make_info(make_alignment(R, r));
        if (pad_quality) ai->quality_scores.append((size_t)pad_quality, '\0');
        algnvector.emplace_back(ai);
    }
    const int n_reads = (int)algnvector.size();
    std::vector<long double> log_likelihood_vec((size_t)S.nbpaths, 0), empty_vec((size_t)S.nbpaths, 0), final_vec((size_t)S.nbpaths, 0);
    int64_t used = 0;
    const double t_loop = omp_get_wtime();
    #pragma omp parallel for num_threads(n_threads) private(log_likelihood_vec) reduction(+:used)
    for (long unsigned int i = 0; i < (long unsigned int)n_reads; ++i) {
        if (algnvector.at(i)->identity < 1e-10) continue;
        log_likelihood_vec = hc.update((int)i, S.pangenome_map, S.nodevector, algnvector.at(i), empty_vec, S.qscore_vec, S.mappabilities,
                                       S.nbpaths, true, S.is_fasta, S.bep, S.incorrect_mapping_vec, n_reads, S.minid, S.is_fasta,
                                       n_threads, graph);
        #pragma omp critical
        for (int j = 0; j < S.nbpaths; ++j) final_vec[(size_t)j] += log_likelihood_vec[(size_t)j];
        used += 1;
    }
    const int maxh_index = (int)(std::max_element(final_vec.begin(), final_vec.end()) - final_vec.begin());
    if (loop_seconds) *loop_seconds = omp_get_wtime() - t_loop;   // the hot loop + max_element only (tables and reads were built before)
    for (int j = 0; j < S.nbpaths; ++j) final_out[j] = final_vec[(size_t)j];
    *argmax_out = maxh_index;
    *used_out = used;
    for (auto* ai : algnvector) delete ai;
    return 0;
}

// Haplocart::get_posterior (src/get_posterior.cpp:87-127) run as it is: parents / children maps in the shape
// load_parents / load_children give them (src/load.cpp:303-345: first token = name, tokens containing '['
// dropped, blank lines skipped -- restated by the caller-side parser below because load.cpp cannot be compiled
// without the euka header tree), result captured from the file write_posterior_log appends to (6 significant
// digits).  Full-precision confidences come from calling the same public member functions
// (sum_log_likelihoods, get_posterior_of_clade) in get_posterior's order.
static std::map<std::string, std::vector<std::string>> parse_rel(const char* txt) {
    std::map<std::string, std::vector<std::string>> out;
    std::istringstream in(txt);
    std::string line;
    while (getline(in, line)) {
        const std::vector<std::string> tokens = allTokensWhiteSpaces(line);
        if (tokens.size() == 0) continue;
        std::vector<std::string> v;
        for (size_t j = 1; j < tokens.size(); ++j) if (tokens[j].find('[') == std::string::npos) v.emplace_back(tokens[j]);
        out.insert(std::make_pair(tokens[0], v));
    }
    return out;
}

int ref_hc_posterior(const long double* final_vec_in, int32_t n_paths, const char* path_names_txt, const char* parents_txt,
                     const char* children_txt, int32_t predicted_idx, const char* tmp_file, int32_t cap, char* clades_out, int32_t clades_cap,
                     double* conf_text_out, int32_t* n_text_out, double* conf_full_out, long double* total_out) {
    Haplocart hc;
    std::vector<std::string> path_names;
    { std::istringstream in(path_names_txt); std::string s; while (getline(in, s)) path_names.push_back(s); }
    if ((int)path_names.size() != n_paths) return -2;
    std::vector<long double> final_vec(final_vec_in, final_vec_in + n_paths);
    std::map<std::string, std::vector<std::string>> parents = parse_rel(parents_txt), children = parse_rel(children_txt);
    // the reference's children.find(p)->second dereferences end() for an absent name (src/get_posterior.cpp:41);
    // refuse such input instead of running undefined behaviour
    const std::string predicted = path_names[(size_t)predicted_idx];
    std::vector<std::string> parent_vec;
    if (parents.count(predicted) > 0) parent_vec = parents.find(predicted)->second;
    // The reference's get_children dereferences children.find(p) unchecked (src/get_posterior.cpp:41): a name that is
    // absent from `children` is undefined behaviour there.  The real children.txt lists every haplogroup (5438 entries,
    // src/test.cpp:617); for test trees that list only inner nodes, every name reachable from an ancestor gets an
    // explicit empty entry ("no children", the meaning the oracle documents) before the reference's code runs.
    {
        std::set<std::string> seen, frontier(parent_vec.begin(), parent_vec.end());
        while (!frontier.empty()) {
            std::set<std::string> next;
            for (const auto& p : frontier) {
                if (!seen.insert(p).second) continue;
                auto it = children.find(p);
                if (it == children.end()) { children.insert(std::make_pair(p, std::vector<std::string>())); continue; }
                for (const auto& c : it->second) next.insert(c);
            }
            frontier.swap(next);
        }
    }

    // full precision: the statements of get_posterior (src/get_posterior.cpp:93-123) on the public helpers
    const long double total_ll = hc.sum_log_likelihoods(final_vec);
    *total_out = total_ll;
    std::vector<long double> all_top{final_vec[(size_t)predicted_idx]};
    int m = 0;
    conf_full_out[m++] = (double)exp(hc.sum_log_likelihoods(all_top) - total_ll);
    all_top.clear();
    std::set<std::string> pred;
    for (size_t j = 0; j < parent_vec.size(); ++j) {
        pred.insert(parent_vec[j]);
        all_top = hc.get_posterior_of_clade(all_top, final_vec, pred, children, total_ll, path_names, false);
        if (all_top.empty()) return -6;   // the reference would read log_lik_vec[0] of an empty vector
        const long double considered = hc.sum_log_likelihoods(all_top);
        // Q8: the reference compares with parent_vec[j-1] at j == 0 (index -1, undefined); "different" here
        if (j == 0 || parent_vec[j] != parent_vec[j - 1]) { if (m < cap) conf_full_out[m] = (double)exp(considered - total_ll); ++m; }
        all_top.clear();
        pred.clear();
    }
    if (m > cap) return -5;

    // and the function itself, end to end, captured from the file it appends to
    // (in a forked child: at j == 0 the function reads parent_vec[-1], Q8 -- whatever the heap holds there must
    // not be able to take the test process down)
    remove(tmp_file);
    {
        const pid_t pid = fork();
        if (pid == 0) {
            hc.get_posterior(final_vec, path_names, parents, children, "sample", predicted, tmp_file, false, 0);
            _exit(0);
        }
        int status = 0;
        if (pid < 0 || waitpid(pid, &status, 0) != pid || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
            // the function itself died on its own undefined read; the full-precision values above stand
            *n_text_out = -7;
            clades_out[0] = 0;
            remove(tmp_file);
            return m;
        }
    }
    std::ifstream f(tmp_file);
    std::string line, payload;
    while (getline(f, line)) if (line.compare(0, 7, "sample\t") == 0) payload = line;
    const std::vector<std::string> tok = allTokens(payload, '\t');
    int n = 0;
    std::string names;
    for (size_t k = 1; k + 1 < tok.size(); k += 3) {
        if (n >= cap) return -5;
        names += tok[k]; names += '\n';
        conf_text_out[n] = atof(tok[k + 1].c_str());
        ++n;
    }
    if ((int)names.size() + 1 > clades_cap) return -5;
    memcpy(clades_out, names.c_str(), names.size() + 1);
    remove(tmp_file);
    *n_text_out = n;
    return m;
}

// ---- table functions, the reference's own (src/miscfunc.h:180-216, src/haplocart_functions.cpp:81-107,
//      src/get_p_obs_base.cpp:38-69) ----
void ref_hc_qscore_vec(double* out100) { const std::vector<double> v = get_qscore_vec(); for (int i = 0; i < 100; ++i) out100[i] = v[(size_t)i]; }
void ref_hc_incorrect_mapping_vec(double* out100) { Haplocart hc; const std::vector<double> v = hc.precompute_incorrect_mapping_probs(); for (int i = 0; i < 100; ++i) out100[i] = v[(size_t)i]; }
double ref_hc_p_seq_error(int Q) { return get_p_seq_error(Q); }
double ref_hc_background_freq(char b) { Haplocart hc; return hc.get_background_freq(b); }
// get_p_obs_base with epsilon = 0 is the per-region match probability pow(1 - 30 mu, 8)
double ref_hc_region_match_prob(int pangenome_base) { Haplocart hc; return (double)hc.get_p_obs_base(pangenome_base, 'A', 'A', 0.0, 8); }
long double ref_hc_p_obs_base(int pangenome_base, char mapping_base, char graph_base, double epsilon) { Haplocart hc; return hc.get_p_obs_base(pangenome_base, mapping_base, graph_base, epsilon, 8); }

// ---- the product's facade against the reference's own AlignmentInfo ----
// vgan_b200::FlatAlignments::assign (vgan_b200/csrc/host/vgan_haplocart.hpp) is a template over "vector of AlignmentInfo*"; here it
// is instantiated with the REFERENCE'S class (src/AlignmentInfo.h, filled as the readGAM lambda fills it, src/readGAM.h:41-46) and
// its flat arrays are handed back so that a test can hold them against what went in.  Sizes first (all outputs may be NULL).
int ref_facade_flatten(const orc_reads* R, const uint8_t* paired, int64_t* counts5, int64_t* seq_off, char* seq, int64_t* qual_off, char* qual,
                       int32_t* mapq, double* identity, uint8_t* is_paired, int64_t* map_off, int64_t* map_node, int64_t* map_offset,
                       uint8_t* map_rev, int64_t* edit_off, int32_t* edit_from, int32_t* edit_to, int64_t* edit_seq_off, char* edit_seq) {
    std::vector<AlignmentInfo*> v;
    for (int64_t r = 0; r < R->n_reads; ++r) v.emplace_back(make_info(make_alignment(R, r), paired && paired[r]));
    vgan_b200::FlatAlignments F;
    F.assign(v);
    for (auto* ai : v) delete ai;
    const vgh_alignments a = F.view();
    if (counts5) {
        counts5[0] = a.n_reads; counts5[1] = (int64_t)F.map_node.size(); counts5[2] = (int64_t)F.edit_from.size();
        counts5[3] = (int64_t)F.seq.size(); counts5[4] = (int64_t)F.edit_seq.size();
    }
    if (!seq_off) return 0;
    const size_t n = (size_t)a.n_reads, nm = F.map_node.size(), ne = F.edit_from.size();
    memcpy(seq_off, a.seq_off, (n + 1) * 8); memcpy(seq, a.seq, F.seq.size());
    memcpy(qual_off, a.qual_off, (n + 1) * 8); memcpy(qual, a.qual, F.qual.size());
    memcpy(mapq, a.mapq, n * 4); memcpy(identity, a.identity, n * 8); memcpy(is_paired, a.paired, n);
    memcpy(map_off, a.map_off, (n + 1) * 8); memcpy(map_node, a.map_node, nm * 8); memcpy(map_offset, a.map_offset, nm * 8);
    memcpy(map_rev, a.map_rev, nm);
    memcpy(edit_off, a.edit_off, (nm + 1) * 8); memcpy(edit_from, a.edit_from, ne * 4); memcpy(edit_to, a.edit_to, ne * 4);
    memcpy(edit_seq_off, a.edit_seq_off, (ne + 1) * 8); memcpy(edit_seq, a.edit_seq, F.edit_seq.size());
    return 0;
}

// ---- Dup_Remover::remove_duplicates_internal (src/rmdup.cpp:68-110), the reference's own ----
// The paired branch indexes mapping()[n_mappings] (src/rmdup.cpp:50,55), one past the end: undefined in the
// reference, an exception in the shim's container; reported as -3.
int ref_dedup(const orc_reads* R, const uint8_t* paired, uint8_t* is_dup_out) {
    std::vector<AlignmentInfo*> v;
    for (int64_t r = 0; r < R->n_reads; ++r) v.emplace_back(make_info(make_alignment(R, r), paired && paired[r]));
    int rc = 0;
    try {
        Dup_Remover d;
        const auto pr = d.remove_duplicates_internal(&v, 1, true);
        for (int64_t r = 0; r < R->n_reads; ++r) is_dup_out[r] = pr.second->at((size_t)r) ? 1 : 0;
        pr.first->clear();   // the function-local static vector (src/rmdup.cpp:73) keeps growing otherwise
        delete pr.second;
    } catch (const std::out_of_range&) { rc = -3; }
    for (auto* ai : v) delete ai;
    return rc;
}

// ---- Damage::initDeamProbabilities (src/damage.cpp:41-323), the reference's own: the 4x4 matrix for fragment
//      length L at position l, as doubles ----
struct RefDamage { Damage d; };
void* ref_damage_create(const char* deam5p, const char* deam3p) {
    try {
        RefDamage* r = new RefDamage();
        r->d.initDeamProbabilities(deam5p ? deam5p : "", deam3p ? deam3p : "");
        return r;
    } catch (...) { return nullptr; }
}
void ref_damage_destroy(void* h) { delete (RefDamage*)h; }
int ref_damage_matrix(void* h, uint32_t L, uint32_t l, double* out16) {
    RefDamage* r = (RefDamage*)h;
    if (L >= r->d.subDeamDiNuc.size() || l >= r->d.subDeamDiNuc[L].size()) return -6;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) out16[a * 4 + b] = (double)r->d.subDeamDiNuc[L][l].p[a][b];
    return 0;
}
int ref_damage_rows(void* h, uint32_t L) { RefDamage* r = (RefDamage*)h; return L < r->d.subDeamDiNuc.size() ? (int)r->d.subDeamDiNuc[L].size() : -1; }

}  // extern "C"

// =====================================================================================
// euka: the per-read lambda of readGAM3 (src/readGAM_Euka.h:67-577), the reference's own, run over injected
// alignments (the shim's vg::io::for_each replays ref_shim::injected()).  Clade / bin tables come through the
// reference's own loaders (Euka::load_clade_info / load_clade_chunks, src/load.cpp:71-157) from the files named;
// the damage tables through Damage::initDeamProbabilities.  The model tables base_freq / t_T_ratio / rare_bases are
// filled with the assignments of Euka::run (src/Euka.cpp:444-486) into zero-initialised arrays (the reference leaves
// the other entries uninitialised and declares rare_bases one entry too short for 'Y'; 256 entries here).
// Per-read results are recovered from the per-clade vectors in arrival order.
// =====================================================================================
extern "C" {

int ref_euka_run(const char* clade_file, const char* bins_file, const char* deam5p_file, const char* deam3p_file,
                 const orc_graph* g, const orc_reads* R, int pad_quality, uint32_t minimum_mq, int32_t n_clades,
                 int32_t* clade_out, double* like_out, double* not_like_out, uint8_t* pass_out, int64_t* clade_count,
                 double* bin_cov, int64_t bin_cov_cap) {
    try {
        Euka ek;
        const int lengthToProf = 5;
        vector<Clade*>* clade_vec = ek.load_clade_info(clade_file, lengthToProf);
        if ((int)clade_vec->size() != n_clades * 6) return -2;
        vector<vector<tuple<int, int, double, double>>> chunks = ek.load_clade_chunks(bins_file);
        Damage dmg;
        dmg.initDeamProbabilities(deam5p_file ? deam5p_file : "", deam3p_file ? deam3p_file : "");
        const vector<double> qscore_vec = ek.get_qscore_vec();
        static double base_freq[256];
        static double t_T_ratio['T' + 1]['T' + 1];
        static bool rare_bases[256];
        memset(base_freq, 0, sizeof(base_freq)); memset(t_T_ratio, 0, sizeof(t_T_ratio)); memset(rare_bases, 0, sizeof(rare_bases));
        base_freq['A'] = log(0.362815); base_freq['C'] = log(0.207743); base_freq['G'] = log(0.116809);
        base_freq['N'] = log(0.25); base_freq['T'] = log(0.312435);
        t_T_ratio['A']['A'] = 1.0; t_T_ratio['C']['C'] = 1.0; t_T_ratio['G']['G'] = 1.0; t_T_ratio['T']['T'] = 1.0;
        t_T_ratio['A']['C'] = 0.02381; t_T_ratio['A']['G'] = 0.95238; t_T_ratio['A']['T'] = 0.02381;
        t_T_ratio['C']['A'] = 0.02381; t_T_ratio['C']['G'] = 0.02381; t_T_ratio['C']['T'] = 0.95238;
        t_T_ratio['G']['A'] = 0.95238; t_T_ratio['G']['C'] = 0.02381; t_T_ratio['G']['T'] = 0.02381;
        t_T_ratio['T']['A'] = 0.02381; t_T_ratio['T']['C'] = 0.95238; t_T_ratio['T']['G'] = 0.02381;
        for (const char c : std::string("WMKRYBDHV")) rare_bases[(int)c] = true;

        const bdsg::ODGI graph = make_graph(g);
        std::vector<NodeInfo*> nodevector;  // passed through, never read by the lambda
        auto& inj = ref_shim::injected();
        inj.clear();
        for (int64_t r = 0; r < R->n_reads; ++r) inj.emplace_back(make_alignment(R, r, pad_quality));
        const unsigned int MINIMUMMQ = minimum_mq, MINNUMOFREADS = 10, MINNUMOFBINS = 6;
        const double ENTROPY = 1.17;
        const int MAXBINS = 0;
        // per read: which clade's vectors grew, by how much count grew
        // (run once over everything; then walk the reads again to attribute the per-clade entries in order)
        auto res = readGAM3(graph, clade_file, false, clade_vec, nodevector, qscore_vec, base_freq, t_T_ratio, rare_bases, chunks,
                            dmg.subDeamDiNuc, lengthToProf, "", MINIMUMMQ, MINNUMOFREADS, MINNUMOFBINS, ENTROPY, MAXBINS);
        inj.clear();
        // attribution: the clade of a read is the LAST bin over all clades containing its first node id (:115-140);
        // recomputed here only to know WHICH clade vector holds the read's entry -- the values are the reference's
        std::vector<size_t> next((size_t)n_clades, 1);  // index 0 is the dummy 0.0 each vector is created with (load.cpp:128)
        for (int64_t r = 0; r < R->n_reads; ++r) {
            clade_out[r] = -1; like_out[r] = not_like_out[r] = 0; pass_out[r] = 0;
            if (R->identity[r] == 0) continue;
            const int n_index = (int)R->map_node[R->map_off[r]];
            int c_n = 0;
            for (size_t i = 0; i < chunks.size(); ++i)
                for (size_t j = 0; j < chunks[i].size(); ++j)
                    if (n_index == std::clamp(n_index, get<0>(chunks[i][j]), get<1>(chunks[i][j]))) c_n = (int)i;
            Clade* c = clade_vec->at((size_t)c_n * 6 + 1);
            const size_t k = next[(size_t)c_n]++;
            if (k >= c->clade_like.size() || k >= c->clade_not_like.size()) return -3;
            clade_out[r] = c_n;
            like_out[r] = c->clade_like[k];
            not_like_out[r] = c->clade_not_like[k];
        }
        for (int c = 0; c < n_clades; ++c) {
            if (next[(size_t)c] != clade_vec->at((size_t)c * 6 + 1)->clade_like.size()) return -4;
            clade_count[c] = clade_vec->at((size_t)c * 6 + 1)->count;
        }
        int64_t b = 0;
        for (size_t i = 0; i < chunks.size(); ++i)
            for (size_t j = 0; j < chunks[i].size(); ++j) { if (b >= bin_cov_cap) return -5; bin_cov[b++] = get<3>(chunks[i][j]); }
        for (auto* c : *clade_vec) delete c;
        delete clade_vec;
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_euka_run: %s\n", e.what()); return -1; }
}

// MCMC::get_proposal_likelihood (src/MCMC.cpp:1175-1213) is private to a chain set-up; not exported here.

// =====================================================================================
// soibean: analyse_GAM (src/getLCAfromGAM.h:31-700), the reference's own, over injected alignments.
// `nodepaths` is built as soibean::run builds it (src/soibean.cpp:476-491): for node minid + k the names of the paths
// through it -- here from the membership matrix instead of ODGI steps.  The R x P matrix is read back from each read's
// pathMap by path name (the reference's convertMapsToVector walks the unordered_map in its iteration order,
// src/soibean.cpp:131-155; by-name lookup gives the same values in pathNames order).
// =====================================================================================
int ref_soibean_run(const orc_graph* g, const orc_reads* R, const uint8_t* member, int64_t member_rows, int32_t member_min_id,
                    int32_t n_paths, int32_t penalty, const char* deam5p_file, const char* deam3p_file, int pad_quality, double* out,
                    uint8_t* used) {
    try {
        Euka ek;
        Damage dmg;
        dmg.initDeamProbabilities(deam5p_file ? deam5p_file : "", deam3p_file ? deam3p_file : "");
        const vector<double> qscore_vec = ek.get_qscore_vec();   // soibean uses Euka's table (src/soibean.cpp:471)
        const bdsg::ODGI graph = make_graph(g);
        std::vector<std::string> pathNames;
        for (int p = 0; p < n_paths; ++p) { char b[32]; snprintf(b, sizeof(b), "path%06d", p); pathNames.emplace_back(b); }
        const int minid = member_min_id;
        std::vector<NodeInfo*> nodevector;
        std::vector<std::vector<std::string>> nodepaths;
        for (int64_t k = 0; k < member_rows; ++k) {
            nodevector.emplace_back(new NodeInfo(minid + k, 0, 0));
            std::vector<std::string> paths;
            for (int p = 0; p < n_paths; ++p) if (member[(size_t)k * (size_t)n_paths + (size_t)p]) paths.emplace_back(pathNames[(size_t)p]);
            nodepaths.push_back(paths);
        }
        std::vector<Clade*> clade_vec;
        auto& inj = ref_shim::injected();
        inj.clear();
        for (int64_t r = 0; r < R->n_reads; ++r) inj.emplace_back(make_alignment(R, r, pad_quality));
        std::streambuf* old = std::cerr.rdbuf();
        std::ostringstream sink;
        std::cerr.rdbuf(sink.rdbuf());   // the function prints the GAM file name
        vector<AlignmentInfo*>* gam = nullptr;
        try {
            gam = analyse_GAM(graph, "/dev/null", &clade_vec, nodevector, nodepaths, pathNames, qscore_vec, false, minid, n_paths, penalty,
                              false, "", false, dmg.subDeamDiNuc);
        } catch (...) { std::cerr.rdbuf(old); inj.clear(); throw; }
        std::cerr.rdbuf(old);
        inj.clear();
        size_t k = 0;
        for (int64_t r = 0; r < R->n_reads; ++r) {
            used[r] = 0;
            for (int p = 0; p < n_paths; ++p) out[(size_t)r * (size_t)n_paths + (size_t)p] = 0;
            if (R->identity[r] == 0) continue;
            if (k >= gam->size()) return -3;
            AlignmentInfo* ai = gam->at(k++);
            used[r] = 1;
            for (int p = 0; p < n_paths; ++p) out[(size_t)r * (size_t)n_paths + (size_t)p] = ai->pathMap.at(pathNames[(size_t)p]);
        }
        if (k != gam->size()) return -4;
        for (auto* ai : *gam) delete ai;
        delete gam;
        for (auto* n : nodevector) delete n;
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_soibean_run: %s\n", e.what()); return -1; }
}

// MCMC::computeBaseLogLike (src/MCMC.h:111-296), the reference's own, for one supported column.
int ref_soibean_base_loglike(char read_base, char graph_base, double ll, const double* freqs7, double t, double branch_len, double* out) {
    try {
        MCMC mc;
        AlignmentInfo::BaseInfo bi;
        bi.readBase = read_base; bi.referenceBase = graph_base; bi.pathSupport = true; bi.logLikelihood = ll;
        RunTreeProportionParams params(nullptr);
        params.freqs['A'] = freqs7[0]; params.freqs['C'] = freqs7[1]; params.freqs['G'] = freqs7[2]; params.freqs['T'] = freqs7[3];
        params.freqs['R'] = freqs7[4]; params.freqs['Y'] = freqs7[5]; params.freqs['M'] = freqs7[6];
        AlignmentInfo read;
        read.detailMap["p"] = std::vector<std::vector<AlignmentInfo::BaseInfo>>{{bi}};
        *out = mc.computeBaseLogLike(&read, params, 0, 0, "p", t, branch_len);
        return 0;
    } catch (...) { return -1; }
}

}  // extern "C"
