// Loader for ViennaRNA "RNAfold parameter file v2.0" energy tables (host side of the product).
#include "params.hpp"

#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace sfb {
namespace {

struct Section {
    std::vector<std::string> tokens;
};

// Splits the file into "# name" sections of whitespace separated tokens, dropping /* */ comments.
std::map<std::string, Section> read_sections(const std::string &path, bool &besteffort) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open parameter file: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    if (text.compare(0, 30, "## RNAfold parameter file v2.0") != 0)
        throw std::runtime_error(path + ": not an 'RNAfold parameter file v2.0'");
    besteffort = text.find("besteffort=1") != std::string::npos;
    std::string clean;
    clean.reserve(text.size());
    for (size_t i = 0; i < text.size();) {
        if (text[i] == '/' && i + 1 < text.size() && text[i + 1] == '*') {
            size_t e = text.find("*/", i + 2);
            i = (e == std::string::npos) ? text.size() : e + 2;
            clean.push_back(' ');
        } else {
            clean.push_back(text[i++]);
        }
    }
    std::map<std::string, Section> out;
    std::istringstream in(clean);
    std::string line, current;
    while (std::getline(in, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos) continue;
        if (line[p] == '#') {
            size_t q = line.find_first_not_of("# \t", p);
            if (q == std::string::npos) {
                current.clear();
                continue;
            }
            std::istringstream ls(line.substr(q));
            ls >> current;
            if (current == "END") current.clear();
            continue;
        }
        if (current.empty()) continue;
        std::istringstream ls(line);
        std::string tok;
        while (ls >> tok) out[current].tokens.push_back(tok);
    }
    return out;
}

struct Cursor {
    const std::vector<std::string> *t;
    size_t pos = 0;
    std::string name;
    int next() {
        if (pos >= t->size()) throw std::runtime_error("parameter block '" + name + "' is too short");
        const std::string &s = (*t)[pos++];
        if (s == "INF") return INF;
        if (s == "DEF") return -50;
        return (int)std::strtol(s.c_str(), nullptr, 10);
    }
};

Cursor open(const std::map<std::string, Section> &secs, const std::string &name) {
    auto it = secs.find(name);
    if (it == secs.end()) throw std::runtime_error("parameter block '" + name + "' missing");
    Cursor c;
    c.t = &it->second.tokens;
    c.name = name;
    return c;
}

void read_mismatch(const std::map<std::string, Section> &secs, const std::string &name, int dst[8][5][5]) {
    Cursor c = open(secs, name);
    std::memset(dst, 0, sizeof(int) * 8 * 25);
    for (int t = 1; t <= 7; t++)
        for (int a = 0; a < 5; a++)
            for (int b = 0; b < 5; b++) dst[t][a][b] = c.next();
}

int loop_key(const std::string &s) {
    int key = 0, mul = 1;
    for (char ch : s) {
        key += encode_nt((unsigned char)ch) * mul;
        mul *= 5;
    }
    return key;
}

inline int imin(int a, int b) { return a < b ? a : b; }

// one set of tables: suffix "" = free energies at 37 C, "_enthalpies" = enthalpies; col = 0 / 1 selects the value / its
// enthalpy in the blocks that interleave them (ML_params, NINIO, Misc, special loops).  Multi / exterior mismatches and
// dangles are stored raw (unclipped).
void read_set(const std::map<std::string, Section> &secs, const std::string &suf, int col, MfeTables &m, double *lxc) {
    std::memset(&m, 0, sizeof m);
    {
        Cursor c = open(secs, "stack" + suf);
        for (int a = 1; a <= 7; a++)
            for (int b = 1; b <= 7; b++) m.stack[a][b] = c.next();
    }
    read_mismatch(secs, "mismatch_hairpin" + suf, m.mismatchH);
    read_mismatch(secs, "mismatch_interior" + suf, m.mismatchI);
    read_mismatch(secs, "mismatch_interior_1n" + suf, m.mismatch1nI);
    read_mismatch(secs, "mismatch_interior_23" + suf, m.mismatch23I);
    read_mismatch(secs, "mismatch_multi" + suf, m.mismatchM);
    read_mismatch(secs, "mismatch_exterior" + suf, m.mismatchExt);
    {
        Cursor c = open(secs, "dangle5" + suf);
        for (int t = 1; t <= 7; t++)
            for (int a = 0; a < 5; a++) m.dangle5[t][a] = c.next();
    }
    {
        Cursor c = open(secs, "dangle3" + suf);
        for (int t = 1; t <= 7; t++)
            for (int a = 0; a < 5; a++) m.dangle3[t][a] = c.next();
    }
    {
        Cursor c = open(secs, "int11" + suf);
        for (int t1 = 1; t1 <= 7; t1++)
            for (int t2 = 1; t2 <= 7; t2++)
                for (int a = 0; a < 5; a++)
                    for (int b = 0; b < 5; b++) m.int11[t1][t2][a][b] = c.next();
    }
    {
        Cursor c = open(secs, "int21" + suf);
        for (int t1 = 1; t1 <= 7; t1++)
            for (int t2 = 1; t2 <= 7; t2++)
                for (int a = 0; a < 5; a++)
                    for (int b = 0; b < 5; b++)
                        for (int d = 0; d < 5; d++) m.int21[t1][t2][a][b][d] = c.next();
    }
    {
        Cursor c = open(secs, "int22" + suf);
        for (int t1 = 1; t1 <= 6; t1++)
            for (int t2 = 1; t2 <= 6; t2++)
                for (int a = 1; a <= 4; a++)
                    for (int b = 1; b <= 4; b++)
                        for (int d = 1; d <= 4; d++)
                            for (int e = 1; e <= 4; e++) m.int22[t1][t2][a][b][d][e] = c.next();
        // positions holding N (code 0) take the least favourable of the four nucleotides
        for (int t1 = 1; t1 <= 6; t1++)
            for (int t2 = 1; t2 <= 6; t2++)
                for (int idx = 0; idx < 625; idx++) {
                    int v[4] = {idx / 125, (idx / 25) % 5, (idx / 5) % 5, idx % 5};
                    if (v[0] && v[1] && v[2] && v[3]) continue;
                    int best = -INF;
                    for (int a = v[0] ? v[0] : 1; a <= (v[0] ? v[0] : 4); a++)
                        for (int b = v[1] ? v[1] : 1; b <= (v[1] ? v[1] : 4); b++)
                            for (int d = v[2] ? v[2] : 1; d <= (v[2] ? v[2] : 4); d++)
                                for (int e = v[3] ? v[3] : 1; e <= (v[3] ? v[3] : 4); e++)
                                    if (m.int22[t1][t2][a][b][d][e] > best) best = m.int22[t1][t2][a][b][d][e];
                    m.int22[t1][t2][v[0]][v[1]][v[2]][v[3]] = best;
                }
    }
    {
        Cursor c = open(secs, "hairpin" + suf);
        for (int i = 0; i <= 30; i++) m.hairpin[i] = c.next();
    }
    {
        Cursor c = open(secs, "bulge" + suf);
        for (int i = 0; i <= 30; i++) m.bulge[i] = c.next();
    }
    {
        Cursor c = open(secs, "interior" + suf);
        for (int i = 0; i <= 30; i++) m.internal_loop[i] = c.next();
    }
    {
        Cursor c = open(secs, "ML_params");   // cu cu_dH cc cc_dH ci ci_dH
        int v[6];
        for (int i = 0; i < 6; i++) v[i] = c.next();
        m.MLbase = v[0 + col];
        m.MLclosing = v[2 + col];
        m.MLintern = v[4 + col];
    }
    {
        Cursor c = open(secs, "NINIO");       // m m_dH max
        int v[3];
        for (int i = 0; i < 3; i++) v[i] = c.next();
        m.ninio = v[col];
        m.max_ninio = v[2];
    }
    {
        auto it = secs.find("Misc");          // DuplexInit dH TerminalAU dH [lxc lxc_dH]
        if (it == secs.end()) throw std::runtime_error("parameter block 'Misc' missing");
        const auto &t = it->second.tokens;
        if (t.size() < 4) throw std::runtime_error("parameter block 'Misc' is too short");
        m.TerminalAU = (int)std::strtol(t[2 + col].c_str(), nullptr, 10);
        if (lxc) *lxc = t.size() >= 5 ? std::strtod(t[4].c_str(), nullptr) : 107.856;
    }
    auto special = [&](const std::string &name, int &n, int *keys, int *es) {
        n = 0;
        auto it = secs.find(name);
        if (it == secs.end()) return;
        const auto &t = it->second.tokens;
        for (size_t k = 0; k + 2 < t.size() + 1 && k + 1 < t.size(); k += 3) {
            if (n >= MAX_SPECIAL) break;
            keys[n] = loop_key(t[k]);
            const size_t v = k + 1 + col < t.size() ? k + 1 + col : k + 1;
            es[n] = (int)std::strtol(t[v].c_str(), nullptr, 10);
            n++;
        }
    };
    special("Tetraloops", m.n_tetra, m.tetra_key, m.tetra_e);
    special("Triloops", m.n_tri, m.tri_key, m.tri_e);
    special("Hexaloops", m.n_hexa, m.hexa_key, m.hexa_e);
}

}  // namespace

void load_params(const std::string &path, HostParams &hp) {
    bool be = false;
    auto secs = read_sections(path, be);
    hp.besteffort = be;
    hp.path = path;
    read_set(secs, "", 0, hp.g37, &hp.lxc37);
    read_set(secs, "_enthalpies", 1, hp.dH, nullptr);
    set_temperature(hp, 37.0);
}

void set_temperature(HostParams &hp, double T) {
    const MfeTables &g = hp.g37, &h = hp.dH;
    MfeTables &m = hp.mfe;
    m = g;   // keys, counts, max_ninio; every energy is overwritten below unless T == 37
    const bool at37 = std::fabs(T - 37.0) < 1e-9;
    const double tempf = (T + 273.15) / (37.0 + 273.15);
    auto rs = [&](int g37, int dh) -> int {
        if (g37 >= INF) return INF;
        if (at37) return g37;
        return (int)((double)dh - (double)(dh - g37) * tempf);   // truncation toward zero, as the int assignment in C
    };
    auto many = [&](int *dst, const int *a, const int *b, size_t n) {
        for (size_t k = 0; k < n; k++) dst[k] = rs(a[k], b[k]);
    };
    many(&m.stack[0][0], &g.stack[0][0], &h.stack[0][0], 64);
    many(m.hairpin, g.hairpin, h.hairpin, 31);
    many(m.bulge, g.bulge, h.bulge, 31);
    many(m.internal_loop, g.internal_loop, h.internal_loop, 31);
    many(&m.mismatchI[0][0][0], &g.mismatchI[0][0][0], &h.mismatchI[0][0][0], 200);
    many(&m.mismatchH[0][0][0], &g.mismatchH[0][0][0], &h.mismatchH[0][0][0], 200);
    many(&m.mismatch1nI[0][0][0], &g.mismatch1nI[0][0][0], &h.mismatch1nI[0][0][0], 200);
    many(&m.mismatch23I[0][0][0], &g.mismatch23I[0][0][0], &h.mismatch23I[0][0][0], 200);
    many(&hp.mismatchM_raw[0][0][0], &g.mismatchM[0][0][0], &h.mismatchM[0][0][0], 200);
    many(&hp.mismatchExt_raw[0][0][0], &g.mismatchExt[0][0][0], &h.mismatchExt[0][0][0], 200);
    many(&hp.dangle5_raw[0][0], &g.dangle5[0][0], &h.dangle5[0][0], 40);
    many(&hp.dangle3_raw[0][0], &g.dangle3[0][0], &h.dangle3[0][0], 40);
    many(&m.int11[0][0][0][0], &g.int11[0][0][0][0], &h.int11[0][0][0][0], 8 * 8 * 25);
    many(&m.int21[0][0][0][0][0], &g.int21[0][0][0][0][0], &h.int21[0][0][0][0][0], 8 * 8 * 125);
    many(&m.int22[0][0][0][0][0][0], &g.int22[0][0][0][0][0][0], &h.int22[0][0][0][0][0][0], 8 * 8 * 625);
    m.MLbase = rs(g.MLbase, h.MLbase);
    m.MLclosing = rs(g.MLclosing, h.MLclosing);
    m.MLintern = rs(g.MLintern, h.MLintern);
    m.ninio = rs(g.ninio, h.ninio);
    m.max_ninio = g.max_ninio;
    m.TerminalAU = rs(g.TerminalAU, h.TerminalAU);
    many(m.tetra_e, g.tetra_e, h.tetra_e, MAX_SPECIAL);
    many(m.tri_e, g.tri_e, h.tri_e, MAX_SPECIAL);
    many(m.hexa_e, g.hexa_e, h.hexa_e, MAX_SPECIAL);
    hp.lxc = at37 ? hp.lxc37 : hp.lxc37 * tempf;
    hp.temperature = T;
    for (int t = 0; t < 8; t++)
        for (int a = 0; a < 5; a++) {
            m.dangle5[t][a] = imin(0, hp.dangle5_raw[t][a]);
            m.dangle3[t][a] = imin(0, hp.dangle3_raw[t][a]);
            for (int b = 0; b < 5; b++) {
                m.mismatchM[t][a][b] = imin(0, hp.mismatchM_raw[t][a][b]);
                m.mismatchExt[t][a][b] = imin(0, hp.mismatchExt_raw[t][a][b]);
            }
        }
    for (int u = 0; u <= MAX_W; u++)
        m.hairpin_len[u] = (u <= 30) ? m.hairpin[u] : m.hairpin[30] + (int)(hp.lxc * std::log(u / 30.));
}

static double smooth(double X) {  // smoothed clip used for dangles / multi / exterior mismatches
    double x = X / 10.;
    if (x < -1.2283697) return 0;
    if (x > 0.8660254) return X;
    double s = std::sin(x - 0.34242663) + 1;
    return 10. * 0.38490018 * s * s;
}

void make_pf_tables(const HostParams &hp, double T, PfTables &q) {
    const MfeTables &m = hp.mfe;
    std::memset(&q, 0, sizeof q);
    const double kT = (T + 273.15) * 1.98717;
    q.kT = kT;
    auto bf = [kT](int e) { return e >= INF ? 0. : std::exp(-(double)e * 10. / kT); };
    q.pf_scale = std::exp(-(-185 + (T - 37.) * 7.27) / kT);
    if (q.pf_scale < 1) q.pf_scale = 1;
    for (int i = 0; i <= 30; i++) {
        q.expbulge[i] = bf(m.bulge[i]);
        q.expinternal[i] = bf(m.internal_loop[i]);
    }
    for (int i = 0; i <= MAXLOOP; i++) q.expninio[i] = bf(imin(m.max_ninio, i * m.ninio));
    for (int u = 0; u <= MAX_W; u++)
        q.exphairpin_len[u] = (u <= 30) ? bf(m.hairpin[u]) : bf(m.hairpin[30]) * std::exp(-(hp.lxc * std::log(u / 30.)) * 10. / kT);
    q.expMLbase = bf(m.MLbase);
    q.expMLclosing = bf(m.MLclosing);
    q.expMLintern = bf(m.MLintern);
    q.expTermAU = bf(m.TerminalAU);
    for (int k = 0; k < m.n_tetra; k++) q.exptetra[k] = bf(m.tetra_e[k]);
    for (int k = 0; k < m.n_tri; k++) q.exptri[k] = bf(m.tri_e[k]);
    for (int k = 0; k < m.n_hexa; k++) q.exphexa[k] = bf(m.hexa_e[k]);
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) {
            q.expstack[a][b] = bf(m.stack[a][b]);
            for (int k = 0; k < 25; k++) (&q.expint11[a][b][0][0])[k] = bf((&m.int11[a][b][0][0])[k]);
            for (int k = 0; k < 125; k++) (&q.expint21[a][b][0][0][0])[k] = bf((&m.int21[a][b][0][0][0])[k]);
            for (int k = 0; k < 625; k++) (&q.expint22[a][b][0][0][0][0])[k] = bf((&m.int22[a][b][0][0][0][0])[k]);
        }
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 5; b++) {
            q.expdangle5[a][b] = std::exp(smooth(-hp.dangle5_raw[a][b]) * 10. / kT);
            q.expdangle3[a][b] = std::exp(smooth(-hp.dangle3_raw[a][b]) * 10. / kT);
            for (int c = 0; c < 5; c++) {
                q.expmismatchI[a][b][c] = bf(m.mismatchI[a][b][c]);
                q.expmismatchH[a][b][c] = bf(m.mismatchH[a][b][c]);
                q.expmismatch1nI[a][b][c] = bf(m.mismatch1nI[a][b][c]);
                q.expmismatch23I[a][b][c] = bf(m.mismatch23I[a][b][c]);
                q.expmismatchM[a][b][c] = std::exp(smooth(-hp.mismatchM_raw[a][b][c]) * 10. / kT);
                q.expmismatchExt[a][b][c] = std::exp(smooth(-hp.mismatchExt_raw[a][b][c]) * 10. / kT);
            }
        }
}

}  // namespace sfb
