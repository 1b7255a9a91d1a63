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

vg::Alignment make_alignment(const orc_reads* R, int64_t r) {
    vg::Alignment a;
    a.set_sequence(std::string(R->seq + R->seq_off[r], R->seq + R->seq_off[r + 1]));
    a.set_quality(std::string(R->qual + R->qual_off[r], R->qual + R->qual_off[r + 1]));
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
               long double* final_out, int32_t* argmax_out, int64_t* used_out) {
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
