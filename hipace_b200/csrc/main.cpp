// hpb200_run -- command-line driver over the C-ABI, with the calling convention of the reference
// executable (`hipace inputs [key=value ...]`, src/main.cpp + Hipace::Evolve): reads a HiPACE++
// input deck, applies the command-line overrides, runs time steps 0 .. max_step on GPU 0 and
// prints the performance counters of Hipace.cpp:509-553 and the sum|Q| field checksums.
//   hpb200_run --check deck [key=value ...]     parse only (no GPU needed)
#include <hpb200.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
    bool check_only = false;
    int a = 1;
    if (a < argc && !strcmp(argv[a], "--check")) { check_only = true; ++a; }
    if (a >= argc) {
        fprintf(stderr, "usage: %s [--check] <input deck> [key=value ...]\n%s\n", argv[0], hpb_version());
        return 2;
    }
    std::ifstream f(argv[a]);
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[a]); return 2; }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string deck = ss.str();
    std::string ov;
    for (int k = a + 1; k < argc; ++k) {       // key=value  ->  "key = value" line (ParmParse syntax)
        std::string t = argv[k];
        const size_t eq = t.find('=');
        if (eq == std::string::npos) { fprintf(stderr, "override '%s' is not key=value\n", argv[k]); return 2; }
        ov += t.substr(0, eq) + " = " + t.substr(eq + 1) + "\n";
    }
    char summary[8192];
    if (hpb_deck_check(deck.c_str(), ov.c_str(), summary, sizeof summary)) {
        fprintf(stderr, "%s\n", hpb_last_error());
        return 1;
    }
    if (check_only) { printf("%s\n", summary); return 0; }

    int max_step = 0;
    if (const char *p = strstr(summary, "max_step=")) max_step = atoi(p + 9);
    hpb_sim *sim = nullptr;
    if (hpb_sim_create(&sim, deck.c_str(), ov.c_str(), 0)) { fprintf(stderr, "%s\n", hpb_last_error()); return 1; }
    hpb_sim_set_option(sim, "max_step", max_step);
    double slices = 0., ms = 0., pushed = 0.;
    for (int step = 0; step <= max_step; ++step) {
        if (hpb_sim_evolve(sim, step, step, 0)) { fprintf(stderr, "%s\n", hpb_last_error()); return 1; }
        hpb_sim_stats st;
        hpb_sim_get_stats(sim, &st);
        slices += (double)st.n_slices; ms += st.slice_loop_ms; pushed += st.n_plasma_pushed + st.n_beam_pushed;
        printf("step %d: %ld slices in %.3f ms, %.2f multigrid V-cycles / slice, %ld QSA violations\n", step,
               st.n_slices, st.slice_loop_ms, st.n_slices ? (double)st.n_mg_vcycles / st.n_slices : 0.,
               st.n_qsa_violation);
    }
    printf("%.1f slices/s, %.4f ns per particle-step\n", slices / (ms * 1e-3), pushed > 0 ? ms * 1e6 / pushed : 0.);
    const int n = hpb_sim_checksum_count(sim);
    std::vector<double> cs(n > 0 ? n : 1);
    hpb_sim_get_checksums(sim, cs.data());
    printf("checksums of the last step (sum|Q|, tests/checksum/checksumAPI.py):\n");
    for (int k = 0; k < n; ++k) printf("  %-10s %.16e\n", hpb_sim_checksum_name(sim, k), cs[k]);
    hpb_sim_destroy(sim);
    return 0;
}
