// Energy-table model of the fold engine (host side) and its device image.
// Format: ViennaRNA "## RNAfold parameter file v2.0" (SURVEY.md A.3).  The reference loads these tables
// implicitly through `import RNA` (ScanFold.py:37) / `RNA.md()` (ScanFold.py:212).
#pragma once
#include <cstdint>
#include <string>

namespace sfb {

constexpr int INF = 10000000;
constexpr int TURN = 3;
constexpr int MAXLOOP = 30;
constexpr int MAX_SPECIAL = 64;
constexpr int MAX_W = 1024;  // longest fold this build accepts

// Integer (dcal) tables used by the MFE kernels.  Plain-old-data: copied verbatim to the device.
struct MfeTables {
    int stack[8][8];
    int hairpin[31], bulge[31], internal_loop[31];
    int mismatchI[8][5][5], mismatchH[8][5][5], mismatch1nI[8][5][5], mismatch23I[8][5][5];
    int mismatchM[8][5][5], mismatchExt[8][5][5];  // clipped to <= 0 (MFE)
    int dangle5[8][5], dangle3[8][5];              // clipped to <= 0 (MFE)
    int int11[8][8][5][5];
    int int21[8][8][5][5][5];
    int int22[8][8][5][5][5][5];
    int MLbase, MLclosing, MLintern, ninio, max_ninio, TerminalAU;
    int n_tetra, n_tri, n_hexa;
    int tetra_key[MAX_SPECIAL], tetra_e[MAX_SPECIAL];  // key = base-5 number of the loop incl. closing pair
    int tri_key[MAX_SPECIAL], tri_e[MAX_SPECIAL];
    int hexa_key[MAX_SPECIAL], hexa_e[MAX_SPECIAL];
    int hairpin_len[MAX_W + 1];  // hairpin initiation by loop size incl. the lxc*ln(u/30) extrapolation
};

// Boltzmann factors for the partition-function kernels (T fixed at load time).
struct PfTables {
    double kT, pf_scale;
    double expstack[8][8], expbulge[31], expinternal[31], expninio[MAXLOOP + 1];
    double expmismatchI[8][5][5], expmismatchH[8][5][5], expmismatch1nI[8][5][5], expmismatch23I[8][5][5];
    double expmismatchM[8][5][5], expmismatchExt[8][5][5], expdangle5[8][5], expdangle3[8][5];
    double expint11[8][8][5][5];
    double expint21[8][8][5][5][5];
    double expint22[8][8][5][5][5][5];
    double expMLbase, expMLclosing, expMLintern, expTermAU;
    double exptetra[MAX_SPECIAL], exptri[MAX_SPECIAL], exphexa[MAX_SPECIAL];
    double exphairpin_len[MAX_W + 1];
};

struct HostParams {
    MfeTables mfe;   // the working set at `temperature` (multi / exterior mismatches and dangles clipped to <= 0)
    // unclipped copies needed for the smoothed PF factors (at `temperature`)
    int mismatchM_raw[8][5][5], mismatchExt_raw[8][5][5], dangle5_raw[8][5], dangle3_raw[8][5];
    double lxc;      // at `temperature`
    double temperature;
    // what the file holds: free energies at 37 C and enthalpies (every block has an `_enthalpies` twin), raw
    MfeTables g37, dH;
    double lxc37;
    bool besteffort;
    std::string path;
};

// Parses `path` (both the 37 C blocks and their enthalpy twins) and selects 37 C; throws std::runtime_error on failure.
void load_params(const std::string &path, HostParams &out);
// Rescales every table to `temperature_c` the way ViennaRNA's get_scaled_params does (md.temperature, ScanFold.py:213,
// ScanFoldFunctions.py:777): E(T) = dH - (dH - E37) * (T + 273.15) / 310.15 truncated to int, lxc proportional to T.
void set_temperature(HostParams &hp, double temperature_c);
// Builds Boltzmann factors from the working set (call set_temperature first: hp.temperature is the temperature used).
void make_pf_tables(const HostParams &hp, double temperature_c, PfTables &out);

// nucleotide / pair encoding
inline int encode_nt(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 1;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 3;
        case 'U': case 'u': case 'T': case 't': return 4;
        default: return 0;
    }
}

}  // namespace sfb
