// Host-side loader and network pre-processing for xnet_b200.
//
// Host mirror of (XNet tree, /root/reference/source):
//   read_sunet / read_netwinv / read_nuclear_data   xnet_data.F90:285-513
//   net_preprocess                                   xnet_preprocess.F90:505-1011
//   read_reaction_data (iffn pointers)               xnet_data.F90:855-866
//   read_ffn_data_table / _logft                     xnet_ffn.F90:73-137
//   read_thermo_file / read_inab_file                model_input_ascii.F90:5-110, 223-350
// Integer outputs (reaction classification, species->reaction maps) are bit-exact scope; they are
// checked against the shipped net_desc goldens and against the oracle's independent parser.
#include "xb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace xb {
namespace {

const int kNnucr[kNChap] = {2, 3, 4, 3, 4, 5, 6, 4, 5, 6, 5};      // nuclei per chapter
const int kNreactant[kNChap] = {1, 1, 1, 2, 2, 2, 2, 3, 3, 4, 1};  // reactants per chapter

// REACLIB source labels in table order (xnet_preprocess.F90:534-556); position = idesc-1
const char *const kDescKnown[171] = {
    "  ec", "bet+", "betp", "bet-", "betm", "btyk", " bec", "bkmo", "mo92", " ffn", " lmp", "fkth",
    "cf88", "wies", "laur", "bb92", "baka", "rolf", "wawo", "mafo", "ms01", " wag", "bb93", "ca85",
    "ornl", " wfh", "rath", "rpsm", "hg94", "ro95", "wien", "iucf", "smwa", "smmo", "wi2p", "aexp",
    "pexp", "bex+", "bex-", "bqa+", "bhi+", "bec ", "bklq", "nuba", "bede", "bqa+", "gamm", "ja01",
    "nosm", "nacr", "tang", "ta04", "ch04", "ku02", "il04", "go00", "most", "il03", "de04", "da03",
    "vanc", "il01", "ha00", "ku92", "ro04", "sc83", "he98", "po00", "il00", "he95", "sc98", "vo00",
    "il99", "beau", "da04", "ca00", "da77", "bl03", "ch01", "fynb", "De99", "bu96", "ba00", "Ha96",
    "sh03", "tr04", "sc05", "nfis", "GW95", "ch05", "re98", "nb03", "mo03", "wc07", "an06", "ka02",
    "gl07", "GG90", "lg06", "im05", "GG92", "ct07", "ga00", "ww02", "wh87", "dh03", "ha01", "cb01",
    "HF99", "hg95", "pg06", "jm06", "sb05", "SM93", "SM86", "thra", "fy05", "ha04", "pt05", "bk91",
    "bk92", "kg91", "SM91", "iwam", " nen", "nebp", "lznu", "lzan", "ths8", "wc12", "nk06", "mb11",
    "jz10", "cd08", "nw00", "il10", "li10", "chw0", "co10", "kd02", "st08", "pkrF", "br10", "hi09",
    "ld11", "hd08", "ua08", "ol11", "ls09", "ia08", "dc11", "mb07", "wc17", "mo97", "ks03", "nac2",
    "mp17", "cb09", "li12", "ma10", "mm11", "si13", "sa12", "ks12", "hg12", "rk12", "ac12", "gl12",
    "mv09", "lhnn", "lhna"};

// weak-reaction class from the label position (xnet_preprocess.F90:648-669)
int weak_flag_from_label(int idesc, int z_first, int z_second) {
  if (idesc == 1) return 1;
  if (idesc == 10 || idesc == 11) return (z_first > z_second && z_first != 1) ? 3 : 2;
  if ((idesc >= 2 && idesc <= 9) || (idesc >= 38 && idesc <= 45) || (idesc >= 92 && idesc <= 94) ||
      idesc == 130 || idesc == 153 || idesc == 154 || idesc == 157)
    return 4;
  if (idesc == 125 || idesc == 127) return 7;
  if (idesc == 126 || idesc == 128) return 8;
  if (idesc == 170) return 9;
  if (idesc == 171) return 10;
  return 0;
}

std::vector<std::string> read_lines(const std::string &path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("Failed to open input file: " + path);
  std::vector<std::string> out;
  std::string s;
  while (std::getline(f, s)) {
    if (!s.empty() && s.back() == '\r') s.pop_back();
    out.push_back(s);
  }
  return out;
}

std::string field(const std::string &line, size_t pos, size_t len) {  // blank padded fixed field
  std::string r(len, ' ');
  for (size_t i = 0; i < len && pos + i < line.size(); ++i) r[i] = line[pos + i];
  return r;
}

std::string trimmed(const std::string &s) {
  size_t a = s.find_first_not_of(" \t");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t");
  return s.substr(a, b - a + 1);
}

double fortran_real(const std::string &txt) {  // blank -> 0, D exponent accepted
  std::string s = trimmed(txt);
  if (s.empty()) return 0.0;
  for (auto &c : s)
    if (c == 'D' || c == 'd') c = 'E';
  return std::strtod(s.c_str(), nullptr);
}

int fortran_int(const std::string &txt) {
  std::string s = trimmed(txt);
  return s.empty() ? 0 : std::atoi(s.c_str());
}

// list-directed token stream over a range of lines
std::vector<std::string> tokens_from(const std::vector<std::string> &lines, size_t first) {
  std::vector<std::string> tok;
  for (size_t i = first; i < lines.size(); ++i) {
    std::string s = lines[i];
    for (auto &c : s)
      if (c == ',' || c == '\t') c = ' ';
    std::istringstream is(s);
    std::string t;
    while (is >> t) tok.push_back(t);
  }
  return tok;
}

void read_nuclear(const std::string &dir, Network &net) {
  // sunet: one name per line
  std::vector<std::string> sun = read_lines(dir + "/sunet");
  while (!sun.empty() && sun.back().empty()) sun.pop_back();
  // netwinv
  std::vector<std::string> w = read_lines(dir + "/netwinv");
  if (w.size() < 2) throw std::runtime_error("Error reading netwinv file");
  net.ny = fortran_int(field(w[0], 0, 5));
  const int ny = net.ny;
  if ((int)sun.size() != ny) throw std::runtime_error("netwinv /= sunet");
  net.t9i.assign(kNG, 0.0);
  for (int j = 0; j < kNG; ++j) {
    double v = (double)fortran_int(field(w[1], 3 * j, 3));
    net.t9i[j] = (j < kNG - 1) ? 0.01 * v : 0.1 * v;
  }
  net.nname.resize(ny);
  for (int i = 0; i < ny; ++i) {
    net.nname[i] = field(w[2 + i], 0, 5);
    if (trimmed(net.nname[i]) != trimmed(field(sun[i], 0, 5))) throw std::runtime_error("netwinv /= sunet");
  }
  std::vector<std::string> tok = tokens_from(w, 2 + ny);
  if ((int)tok.size() < ny * (6 + kNG)) throw std::runtime_error("Error reading netwinv file");
  net.aa.resize(ny); net.zz.resize(ny); net.nn.resize(ny); net.mex.resize(ny); net.be.resize(ny);
  net.iz.resize(ny); net.in.resize(ny); net.g.resize((size_t)ny * kNG);
  size_t p = 0;
  for (int i = 0; i < ny; ++i) {
    if (trimmed(tok[p]) != trimmed(net.nname[i]))
      throw std::runtime_error("netwinv data entry name does not match header");
    net.aa[i] = fortran_real(tok[p + 1]);
    net.iz[i] = std::atoi(tok[p + 2].c_str());
    net.in[i] = std::atoi(tok[p + 3].c_str());
    net.mex[i] = fortran_real(tok[p + 5]);
    p += 6;
    for (int j = 0; j < kNG; ++j) net.g[(size_t)i * kNG + j] = fortran_real(tok[p + j]);
    p += kNG;
    net.zz[i] = (double)net.iz[i];
    net.nn[i] = (double)net.in[i];
  }
  net.izmax = *std::max_element(net.iz.begin(), net.iz.end());
  // binding energies from the n/p mass excesses of the table itself (xnet_data.F90:486-498)
  const double avn = 6.022140857e+23, epmev = 1.6021766208e-06, clt = 2.99792458e+10;
  const double m_e = 0.5109989461, m_n = 939.5654133, m_p = 938.2720813;
  const double amu = 1.0 / (avn * epmev), m_u = amu * (clt * clt);
  int ineut = -1, iprot = -1;
  for (int i = 0; i < ny; ++i) {
    if (net.iz[i] == 0 && net.in[i] == 1) ineut = i;
    if (net.iz[i] == 1 && net.in[i] == 0) iprot = i;
  }
  net.mex_n = ineut >= 0 ? net.mex[ineut] : m_n - m_u;
  net.mex_p = iprot >= 0 ? net.mex[iprot] : m_p + m_e - m_u;
  for (int i = 0; i < ny; ++i) net.be[i] = net.mex_n * net.nn[i] + net.mex_p * net.zz[i] - net.mex[i];
}

struct RawReaction {
  int32_t n[6];
  double an[6];
  double rc[7];
  double q;
  int32_t iwk, ires, irev;
};

void preprocess_reactions(const std::string &dir, Network &net) {
  const int ny = net.ny;
  std::vector<std::string> L = read_lines(dir + "/netsu");
  if (L.empty()) throw std::runtime_error("Failed to read netsu");
  net.nffn = fortran_int(field(L[0], 0, 5));
  net.nnnu = fortran_int(field(L[0], 5, 6));
  while (L.size() > 1 && trimmed(L.back()).empty()) L.pop_back();

  int cla[kNChap + 1], cle[kNChap + 1];
  for (int j = 0; j <= kNChap; ++j) { cla[j] = 1; cle[j] = 0; }
  std::vector<RawReaction> store[4];  // slot n-1 of the class, overwritten like the Fortran arrays
  int count[4] = {0, 0, 0, 0};
  int n = 1, chapter = 1;

  for (size_t pos = 1; pos + 3 <= L.size(); pos += 3) {
    const std::string &hdr = L[pos];
    const int k = fortran_int(field(hdr, 0, 2));
    if (k != 0) {  // chapter marker: (re)position the running reaction index
      if (k == 1) { chapter = k; cla[k] = 1; n = 1; }
      else if (k == 4 || k == 8 || k == 10) { chapter = k; cla[k] = 1; cle[k - 1] = n - 1; n = 1; }
      else if (k == 11) { chapter = k; cla[k] = cle[3] + 1; cle[k - 1] = n - 1; n = cle[3] + 1; }
      else { chapter = k; cla[k] = n; cle[k - 1] = n - 1; }
      continue;
    }
    std::string name[6];
    int32_t idx[6];
    bool outside = false;
    for (int s = 0; s < 6; ++s) {
      name[s] = field(hdr, 5 + 5 * s, 5);
      idx[s] = index_from_name(net, name[s]);
      if (idx[s] > ny) outside = true;
    }
    if (outside) continue;  // reaction involves a nucleus not in sunet: dropped
    const std::string label = field(hdr, 43, 4);
    const char resflag = field(hdr, 47, 1)[0], revflag = field(hdr, 48, 1)[0];
    RawReaction r{};
    for (int j = 0; j < 4; ++j) r.rc[j] = fortran_real(field(L[pos + 1], 13 * j, 13));
    for (int j = 0; j < 3; ++j) r.rc[4 + j] = fortran_real(field(L[pos + 2], 13 * j, 13));
    int idesc = 172;
    for (int d = 0; d < 171; ++d)
      if (label == kDescKnown[d]) { idesc = d + 1; break; }
    r.iwk = weak_flag_from_label(idesc, idx[0] > 0 ? net.iz[idx[0] - 1] : 0, idx[1] > 0 ? net.iz[idx[1] - 1] : 0);
    r.ires = (resflag == ' ' || resflag == 'w') ? 0 : (resflag == 'n' ? 1 : 2);
    r.irev = (revflag == 'v') ? 1 : 0;
    const int nreact = kNreactant[chapter - 1];
    // reactants count -1, products +1, blanks 0; Q from the table masses
    int ki[6];
    double qv = 0.0;
    for (int s = 0; s < 6; ++s) {
      const bool blank = (name[s] == "     ");
      if (s + 1 > nreact) {
        ki[s] = blank ? 0 : 1;
        if (!blank) { int j = idx[s] - 1; qv = qv + net.be[j] - net.mex_n * net.nn[j] - net.mex_p * net.zz[j]; }
      } else {
        ki[s] = -1;
        int j = idx[s] - 1;
        qv = qv - net.be[j] + net.mex_n * net.nn[j] + net.mex_p * net.zz[j];
      }
    }
    r.q = qv;
    // fold runs of identical names: reactants, then products (xnet_preprocess.F90:728-750)
    {
      int head = 0;
      for (int s = 1; s < nreact; ++s) {
        if (name[s - 1] != name[s]) head = s;
        else { ki[head] += ki[s]; ki[s] = 0; }
      }
      head = nreact;
      for (int s = nreact + 1; s < 6; ++s) {
        if (name[s - 1] != name[s]) head = s;
        else { ki[head] += ki[s]; ki[s] = 0; }
      }
    }
    double fac = 1.0;
    int nslots = 5;
    if (nreact > 1) {
      int lo = 0;
      for (int s = 0; s < nreact; ++s) lo = std::min(lo, ki[s]);
      int f = 1;
      for (int i = 2; i <= -lo; ++i) f *= i;
      fac = 1.0 / (double)f;
      nslots = 6;
    }
    for (int s = 0; s < 6; ++s) {
      r.n[s] = s < nslots ? idx[s] : 0;
      r.an[s] = s < nslots ? (double)ki[s] * fac : 0.0;
    }
    auto &vec = store[nreact - 1];
    if ((int)vec.size() < n) vec.resize(n);
    vec[n - 1] = r;
    ++count[nreact - 1];
    ++n;
  }
  cle[chapter] = n - 1;
  net.chap_la.assign(cla + 1, cla + 1 + kNChap);
  net.chap_le.assign(cle + 1, cle + 1 + kNChap);

  for (int c = 0; c < 4; ++c) {
    const int nr = count[c];
    net.nreac[c] = nr;
    if ((int)store[c].size() < nr) throw std::runtime_error("netsu: inconsistent chapter markers");
    net.ni[c].assign((size_t)nr * 6, 0); net.rc[c].assign((size_t)nr * 7, 0.0); net.q[c].assign(nr, 0.0);
    net.iwk[c].assign(nr, 0); net.ires[c].assign(nr, 0); net.irev[c].assign(nr, 0);
    for (int k = 0; k < nr; ++k) {
      const RawReaction &r = store[c][k];
      for (int s = 0; s < 6; ++s) net.ni[c][(size_t)k * 6 + s] = r.n[s];
      for (int j = 0; j < 7; ++j) net.rc[c][(size_t)k * 7 + j] = r.rc[j];
      net.q[c][k] = r.q; net.iwk[c][k] = r.iwk; net.ires[c][k] = r.ires; net.irev[c][k] = r.irev;
    }
  }

  // species -> reaction registration: species outer, chapters of the class in order, reaction,
  // slot (xnet_preprocess.F90:863-951).  Defines the summation order of ydot and the Jacobian.
  const double tiny = std::numeric_limits<double>::min();
  net.la.assign((size_t)4 * ny, 0); net.le.assign((size_t)4 * ny, 0);
  for (int c = 0; c < 4; ++c) {
    net.mu[c].clear(); net.a[c].clear();
    int cnt = 0;
    for (int i = 1; i <= ny; ++i) {
      net.la[(size_t)c * ny + i - 1] = cnt + 1;
      for (int j = 1; j <= kNChap; ++j) {
        if (kNreactant[j - 1] != c + 1) continue;
        for (int k = cla[j]; k <= cle[j]; ++k) {
          if (k < 1 || k > (int)store[c].size()) continue;
          const RawReaction &r = store[c][k - 1];
          for (int l = 0; l < kNnucr[j - 1]; ++l) {
            if (r.n[l] == i && std::fabs(r.an[l]) > tiny) {
              ++cnt;
              net.a[c].push_back(r.an[l]);
              net.mu[c].push_back(k);
            }
          }
        }
      }
      net.le[(size_t)c * ny + i - 1] = cnt;
    }
    net.nan[c] = cnt;
  }

  // FFN table pointer lives in REACLIB coefficient 1 (xnet_data.F90:857-866)
  net.iffn.assign(net.nreac[0], 0);
  for (int k = 0; k < net.nreac[0]; ++k)
    if (net.iwk[0][k] == 2 || net.iwk[0][k] == 3) net.iffn[k] = (int32_t)std::lrint(net.rc[0][(size_t)k * 7]);
}

}  // namespace (anonymous)

// match_react (xnet_preprocess.F90:108-289) with flux_search (:55-106): every reaction is assigned to a flux set
// keyed by its (reactants | products) written in the exoergic direction; non-weak reactions share the set of their
// reverse, weak reactions (iwk != 0) always open a new one.  Then flux_init (xnet_flux.F90:20-65): double-counting
// factors and the originating / terminating species of each set.
void match_react(Network &net) {
  net.mflx = 0; net.nflx.clear(); net.iwflx.clear(); net.qflx.clear();
  for (int c = 0; c < 4; ++c) net.ifl[c].assign(net.nreac[c], 0);
  if (net.chap_la.size() != (size_t)kNChap) return;   // networks built from arrays carry no chapter extents
  for (int j = 0; j < kNChap; ++j) {
    const int nreact = kNreactant[j], nnuc = kNnucr[j], nprod = nnuc - nreact, c = nreact - 1;
    for (int k = net.chap_la[j]; k <= net.chap_le[j]; ++k) {
      if (k < 1 || k > net.nreac[c]) continue;
      const int32_t *n = &net.ni[c][(size_t)(k - 1) * 6];
      const double q = net.q[c][k - 1];
      const int iw = net.iwk[c][k - 1];
      int32_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const bool fwd = q > 0.0;
      if (fwd) {
        for (int s = 0; s < nreact; ++s) t[s] = n[s];
        for (int s = 0; s < nprod; ++s) t[4 + s] = n[nreact + s];
      } else {
        for (int s = 0; s < nprod; ++s) t[s] = n[nreact + s];
        for (int s = 0; s < nreact; ++s) t[4 + s] = n[s];
      }
      int found = 0;
      if (iw == 0)
        for (int m = 0; m < net.mflx; ++m)
          if (std::equal(t, t + 8, &net.nflx[(size_t)m * 8])) { found = m + 1; break; }
      if (!found) {
        net.nflx.insert(net.nflx.end(), t, t + 8);
        net.qflx.push_back(std::fabs(q));
        net.iwflx.push_back(iw);
        found = ++net.mflx;
      }
      net.ifl[c][k - 1] = fwd ? found : -found;
    }
  }
  flux_init(net);
}

// flux_init (xnet_flux.F90:20-65): double-counting factors and originating / terminating species of every pair
void flux_init(Network &net) {
  const double fact[5] = {1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 4.0};
  net.dcflx.assign((size_t)2 * net.mflx, 1.0); net.ifl_orig.assign(net.mflx, 0); net.ifl_term.assign(net.mflx, 0);
  for (int m = 0; m < net.mflx; ++m) {
    const int32_t *f = &net.nflx[(size_t)m * 8];
    int cf = 1, cr = 1, nf = 0, nr = 0;
    for (int s = 1; s < 4; ++s) if (f[s] != 0 && f[s] == f[s - 1]) ++cf;
    for (int s = 5; s < 8; ++s) if (f[s] != 0 && f[s] == f[s - 1]) ++cr;
    if (cf == 3 && f[1] != f[2]) cf = 5;
    if (cr == 3 && f[5] != f[6]) cr = 5;
    net.dcflx[(size_t)2 * m] = fact[cf - 1]; net.dcflx[(size_t)2 * m + 1] = fact[cr - 1];
    for (int s = 0; s < 4; ++s) { if (f[s] != 0) ++nf; if (f[4 + s] != 0) ++nr; }
    net.ifl_orig[m] = nf ? f[nf - 1] : 0;
    net.ifl_term[m] = nr ? f[4 + nr - 1] : 0;
  }
}

// Structural pattern of the Jacobian from the maps (d ydot_i / d y_j is structurally non-zero iff j is a reactant of a
// reaction registered with species i, xnet_jacobian_dense.F90:311-341), the light-particles-last order, its envelope.
void structured_lu_plan(Network &net) {
  const int ny = net.ny;
  std::vector<int> light, heavy;
  for (int i = 0; i < ny; ++i) ((net.zz[i] <= 2.0 && net.aa[i] <= 4.0) ? light : heavy).push_back(i);
  net.slu_perm.clear();
  for (int i : heavy) net.slu_perm.push_back(i + 1);
  for (int i : light) net.slu_perm.push_back(i + 1);
  std::vector<int> pos(ny);
  for (int p = 0; p < ny; ++p) pos[net.slu_perm[p] - 1] = p;
  std::vector<int> rowfirst(ny), colfirst(ny);
  for (int p = 0; p < ny; ++p) { rowfirst[p] = p; colfirst[p] = p; }
  for (int c = 0; c < 4; ++c) {
    if (net.la.size() != (size_t)4 * ny) break;
    for (int i = 0; i < ny; ++i) {
      for (int t = net.la[(size_t)c * ny + i]; t <= net.le[(size_t)c * ny + i]; ++t) {
        if (t < 1 || t > (int)net.mu[c].size()) continue;
        const int k = net.mu[c][t - 1] - 1;
        for (int r = 0; r <= c; ++r) {
          const int j = net.ni[c][(size_t)k * 6 + r] - 1;
          if (j < 0 || j >= ny) continue;
          const int pi = pos[i], pj = pos[j];
          if (pj < pi) rowfirst[pi] = std::min(rowfirst[pi], pj);
          else if (pi < pj) colfirst[pj] = std::min(colfirst[pj], pi);
        }
      }
    }
  }
  net.slu_rowfirst.resize(ny); net.slu_colfirst.resize(ny);
  net.slu_doubles = ny;
  for (int p = 0; p < ny; ++p) {
    net.slu_rowfirst[p] = rowfirst[p] + 1; net.slu_colfirst[p] = colfirst[p] + 1;
    net.slu_doubles += (p - rowfirst[p]) + (p - colfirst[p]);
  }
}

// Everything the structured-LU kernels index with (see SluTables).  The pattern is the structural one of
// structured_lu_plan (species i depends on every reactant of every reaction registered with it); with self-heating the
// T9 row and column are full.  The light particles and T9 form the pivoted border.
void structured_lu_tables(const Network &net, int heat, SluTables &t) {
  const int ny = net.ny, n = ny + (heat ? 1 : 0);
  t = SluTables();
  t.n = n;
  int nlight = 0;
  for (int i = 0; i < ny; ++i) if (net.zz[i] <= 2.0 && net.aa[i] <= 4.0) ++nlight;
  t.nbord = nlight + (heat ? 1 : 0);
  t.perm.resize(n); t.pos.resize(n);
  for (int p = 0; p < ny; ++p) t.perm[p] = net.slu_perm[p] - 1;
  if (heat) t.perm[ny] = ny;
  for (int p = 0; p < n; ++p) t.pos[t.perm[p]] = p;
  t.rf.resize(n); t.cf.resize(n);
  for (int p = 0; p < ny; ++p) { t.rf[p] = net.slu_rowfirst[p] - 1; t.cf[p] = net.slu_colfirst[p] - 1; }
  if (heat) { t.rf[ny] = 0; t.cf[ny] = 0; }
  // the border rows and columns are stored in full: the pivoting inside the border exchanges whole rows
  for (int p = n - t.nbord; p < n; ++p) { t.rf[p] = 0; t.cf[p] = 0; }
  const int nh = n - t.nbord;
  // Pair alignment (XB_SLU_PAIR=1, off by default): slu_factor then eliminates the pivots (k, k+1), k even, in ONE round (a
  // rank-2 update).  For that the two pivots must work on the same rows and columns: row and column profiles are rounded
  // down to even, i.e. a stored zero at (i, k) for every row whose pattern starts in column k+1 and at (k, j) likewise
  // (SN160: +2 % storage).  Bit-identical to the pivot-by-pivot elimination and parity-green, but measured SLOWER on B200
  // (SN160 + screening 2.44e6 against 2.63e6 zone-it/s): every lane re-forms the multipliers of the second pivot for its
  // rows (+50 % shared-memory loads and FP64 work) and the two reciprocals are a serial chain; see DESIGN 4c.
  if (const char *v = std::getenv("XB_SLU_PAIR")) t.pair = (*v && std::atoi(v) != 0) ? 1 : 0;
  if (t.pair) for (int p = 0; p < nh; ++p) { t.rf[p] &= ~1; t.cf[p] &= ~1; }
  // spans of the heavy part per pivot, last stored heavy row per column
  t.nrb.assign(nh, 0); t.ncb.assign(nh, 0); t.cbot.assign(n, nh - 1);
  for (int k = 0; k < nh; ++k) {
    int rhi = k, chi = k;
    for (int i = k + 1; i < nh; ++i) { if (t.rf[i] <= k) rhi = i; if (t.cf[i] <= k) chi = i; }
    t.nrb[k] = rhi - k; t.ncb[k] = chi - k; t.cbot[k] = rhi;
  }
  t.cb.resize(n);
  // XB_SLU_PADW = 1 (no padding) .. 16 (any 16 consecutive columns: 16 distinct bank pairs).  8 is the measured optimum on
  // SN160: 0.94k doubles less padding than 16 lets all 2048 rates stay cached in shared memory (assembly -7 %, LU +1 %)
  int padw = 8;
  if (const char *v = std::getenv("XB_SLU_PADW")) if (*v) padw = std::min(16, std::max(1, std::atoi(v)));
  int o = 0;
  for (int j = 0; j < n; ++j) {
    // start address a of row cf[j] such that cb[j] = a - cf[j] differs (mod 16: the 64-bit shared-memory bank pair) from
    // the bases of the padw - 1 columns before it
    int a = o;
    auto res = [&](int base) { return ((base % 16) + 16) % 16; };
    for (;; ++a) {
      bool clash = false;
      for (int q = std::max(0, j - (padw - 1)); q < j && !clash; ++q) clash = res(t.cb[q]) == res(a - t.cf[j]);
      if (!clash) break;
    }
    t.cb[j] = a - t.cf[j];
    o = a + (t.cbot[j] - t.cf[j] + 1);
  }
  // border rows, column by column with an odd stride (consecutive columns -> distinct banks); column n carries the
  // right-hand side entries of the border rows through the factorisation
  t.bo = o; t.bs = t.nbord | 1;
  o += (n + 1) * t.bs;
  t.nstore = o;
  for (int k = 0; k < n; ++k) {
    double nr = 0, nc = 0;
    for (int i = k + 1; i < n; ++i) { if (t.rf[i] <= k) nr += 1; if (t.cf[i] <= k) nc += 1; }
    t.flops += nr + 2.0 * nr * nc;
    if (k < nh) { const double er = t.nrb[k] + t.nbord, ec = t.ncb[k] + t.nbord; t.flops_envelope += er + 2.0 * er * ec; }
    else { const double e = n - 1 - k; t.flops_envelope += e + 2.0 * e * e; }
  }
}

void structured_lu_schedule(const SluTables &t, int nwl, int rhs_off, int ntest_rows, SluSchedule &s) {
  const int n = t.n, nb = t.nbord, nh = n - nb, bo = t.bo, bs = t.bs;
  s = SluSchedule();
  s.nwl = nwl;
  // column j = n is the right-hand side: base rhs_off, stored from row 0, border rows in column n of the border block
  auto colbase = [&](int j) { return j < n ? t.cb[j] : rhs_off; };
  auto colfirst = [&](int j) { return j < n ? t.cf[j] : 0; };
  struct Item { int k, q, c0; bool test; };
  auto idle_round = [&](int pv, int k) {
    for (int w = 0; w < nwl; ++w) {
      s.warp.push_back((uint32_t)pv | ((uint32_t)pv << 16));
      s.warp.push_back((uint32_t)k << 8);
      for (int l = 0; l < 32; ++l) s.lane.push_back((uint32_t)pv | ((uint32_t)pv << 16));
    }
    ++s.nround;
  };
  const int nbt = ntest_rows - nh;                  // tested border rows
  for (int k = 0; k < nh; ++k) {
    const int ncb = t.ncb[k], nrb = t.nrb[k], ncx = nb + 1;
    const int nchunk = (nrb + 7) / 8;               // band chunks; chunk number nchunk = the border rows
    const int nslot = (ncb <= 16) ? 16 + ncx : ncb + ncx;
    const int pv = t.cb[k] + k, lband = pv + 1, lbord = bo + k * bs;
    // lane slots: 0..15 the first 16 band columns (16 distinct bank pairs per row), then the border columns and the
    // right-hand side, then the band columns past the 16th
    auto slot_col = [&](int c) {
      if (c < 16) return c < ncb ? k + 1 + c : -1;
      const int e = c - 16;
      if (e < ncx) return nh + e;
      const int cc = e - ncx + 16;
      return cc < ncb ? k + 1 + cc : -1;
    };
    std::vector<Item> items;
    for (int c0 = 0; c0 < nslot; c0 += 32)
      for (int q = 0; q <= nchunk; ++q) {
        bool any = false;
        for (int l = 0; l < 32 && !any; ++l) { const int j = slot_col(c0 + l); any = j >= 0 && colfirst(j) <= k; }
        const bool test = c0 == 0 && (q < nchunk || nbt > 0);
        if (any || test) items.push_back({k, q, c0, test});
      }
    for (size_t i0 = 0; i0 < items.size(); i0 += (size_t)nwl) {
      for (int w = 0; w < nwl; ++w) {
        if (i0 + w >= items.size()) {                 // idle warp
          s.warp.push_back((uint32_t)pv | ((uint32_t)pv << 16));
          s.warp.push_back((uint32_t)k << 8);
          for (int l = 0; l < 32; ++l) s.lane.push_back((uint32_t)pv | ((uint32_t)pv << 16));
          continue;
        }
        const Item &it = items[i0 + w];
        const bool band = it.q < nchunk;
        const int r0 = it.q * 8;
        const int lp = band ? lband + r0 : lbord;
        const int nrow = band ? std::min(8, nrb - r0) : nb;
        const int ntest = it.test ? (band ? nrow : std::max(0, std::min(nbt, nb))) : 0;
        s.warp.push_back((uint32_t)lp | ((uint32_t)pv << 16));
        s.warp.push_back((uint32_t)nrow | ((uint32_t)ntest << 4) | ((uint32_t)k << 8));
        for (int l = 0; l < 32; ++l) {
          const int j = slot_col(it.c0 + l);
          if (j >= 0 && colfirst(j) <= k) {
            const int ap = band ? colbase(j) + k + 1 + r0 : bo + j * bs;
            s.lane.push_back((uint32_t)ap | ((uint32_t)(colbase(j) + k) << 16));
          } else {
            s.lane.push_back((uint32_t)lp | ((uint32_t)pv << 16));
          }
        }
      }
      ++s.nround;
    }
  }
  const int pv_last = nh > 0 ? t.cb[nh - 1] + nh - 1 : 0;
  // the kernel runs three rounds per loop trip and reads three rounds ahead
  while (s.nround % 3) idle_round(pv_last, std::max(nh - 1, 0));
  const int real = s.nround;
  for (int i = 0; i < 3; ++i) idle_round(pv_last, std::max(nh - 1, 0));
  s.nround = real;
}

void structured_lu_schedule2(const SluTables &t, int nwl, int rhs_off, int ntest_rows, SluSchedule2 &s) {
  const int n = t.n, nb = t.nbord, nh = n - nb, bo = t.bo, bs = t.bs;
  s = SluSchedule2();
  s.nwl = nwl;
  auto colbase = [&](int j) { return j < n ? t.cb[j] : rhs_off; };
  auto colfirst = [&](int j) { return j < n ? t.cf[j] : 0; };
  const int nbt = std::max(0, std::min(ntest_rows - nh, nb));          // tested border rows
  int last_pv1 = 0, last_k = 0;
  auto push_idle_warp = [&](int pv1, int u12, int k, bool single) {
    s.warp.push_back((uint32_t)pv1 | ((uint32_t)pv1 << 16));
    s.warp.push_back(single ? (1u << 13) : 0u);
    s.warp.push_back((uint32_t)pv1 | ((uint32_t)u12 << 16));
    s.warp.push_back((uint32_t)k);
    for (int l = 0; l < 32; ++l) { s.lane.push_back((uint32_t)pv1 | ((uint32_t)pv1 << 16)); s.lane.push_back((uint32_t)pv1); }
  };
  for (int k = 0; k < nh; k += 2) {
    const bool single = k + 1 >= nh;
    const int ncb = t.ncb[k], nrb = t.nrb[k], ncx = nb + 1;
    const int row0 = single ? k + 1 : k + 2, nband = std::max(0, single ? nrb : nrb - 1);
    const int nchunk = (nband + 7) / 8;
    const int nslot = (ncb <= 16) ? 16 + ncx : ncb + ncx;
    const int pv1 = t.cb[k] + k, u12 = single ? pv1 : t.cb[k + 1] + k;
    last_pv1 = pv1; last_k = k;
    auto slot_col = [&](int c) {
      if (c < 16) return c < ncb ? k + 1 + c : -1;
      const int e = c - 16;
      if (e < ncx) return nh + e;
      const int cc = e - ncx + 16;
      return cc < ncb ? k + 1 + cc : -1;
    };
    // slot groups last to first, the rows of a group last chunk (the border) to first: the lane of column k+1 (slot 0 of
    // group 0) overwrites the multipliers of pivot k+1 that every other item of the same rows reads, and the item of the
    // first rows hands row k+1 of U over for the deferred store -- both have to come last.  A group starts a new round.
    const int ngrp = (nslot + 31) / 32;
    for (int gi = ngrp - 1; gi >= 0; --gi) {
      const int c0 = gi * 32;
      struct Item { int q; };
      std::vector<Item> items;
      for (int q = nchunk; q >= 0; --q) items.push_back({q});
      for (size_t i0 = 0; i0 < items.size(); i0 += (size_t)nwl) {
        for (int w = 0; w < nwl; ++w) {
          if (i0 + w >= items.size()) { push_idle_warp(pv1, u12, k, single); continue; }
          const int q = items[i0 + w].q;
          const bool band = q < nchunk, lastitem = (i0 + w + 1 == items.size());
          const int r0 = row0 + q * 8;
          const int lp1 = band ? t.cb[k] + r0 : bo + k * bs;
          const int lp2 = single ? lp1 : (band ? t.cb[k + 1] + r0 : bo + (k + 1) * bs);
          const int nrow = band ? std::min(8, nband - q * 8) : nb;
          const int nt = (gi == 0) ? (band ? nrow : nbt) : 0;
          const int commit = (!single && lastitem) ? 1 : 0;
          s.warp.push_back((uint32_t)lp1 | ((uint32_t)lp2 << 16));
          s.warp.push_back((uint32_t)nrow | ((uint32_t)nt << 4) | ((uint32_t)(single ? 0 : nt) << 8) | ((uint32_t)commit << 12) |
                           ((uint32_t)(single ? 1 : 0) << 13));
          s.warp.push_back((uint32_t)pv1 | ((uint32_t)u12 << 16));
          s.warp.push_back((uint32_t)k);
          for (int l = 0; l < 32; ++l) {
            const int j = slot_col(c0 + l);
            if (j >= 0 && colfirst(j) <= k) {
              const int ap = band ? colbase(j) + r0 : bo + j * bs;
              const int up1 = colbase(j) + k, up2 = single ? up1 : up1 + 1;
              const int col2 = (!single && j == k + 1) ? 1 : 0;
              s.lane.push_back((uint32_t)ap | ((uint32_t)up1 << 16));
              s.lane.push_back((uint32_t)up2 | ((uint32_t)col2 << 16));
            } else {
              s.lane.push_back((uint32_t)lp1 | ((uint32_t)pv1 << 16));
              s.lane.push_back((uint32_t)pv1);
            }
          }
        }
        ++s.nround;
      }
    }
  }
  auto idle_round = [&]() { for (int w = 0; w < nwl; ++w) push_idle_warp(last_pv1, last_pv1, last_k, true); ++s.nround; };
  while (s.nround % 3) idle_round();
  const int real = s.nround;
  for (int i = 0; i < 3; ++i) idle_round();
  s.nround = real;
}

void structured_lu_solve_table(const SluTables &T, std::vector<uint32_t> &sl, std::vector<int32_t> &so, int *ngroup_out) {
  const int nh = T.n - T.nbord;
  const int ngroup = ((nh + 3) / 4 + 1) / 2 * 2;
  sl.assign((size_t)(ngroup + 2) * 32 * 2, 0xffffffffu);
  so.clear();
  for (int k = nh - 1; k >= 0; --k) {
    const int stp = nh - 1 - k, g = stp / 4, sidx = stp % 4;
    if (T.cf[k] <= k - 32) so.push_back(k);
    for (int l = 0; l < 32; ++l) {
      const int i = k - ((k - l) & 31);
      unsigned ad = 0xffffu;
      if (i == k || (i >= 0 && i >= T.cf[k])) ad = (unsigned)(T.cb[k] + i);
      uint32_t &wd = sl[((size_t)g * 32 + l) * 2 + sidx / 2];
      wd = (sidx % 2) ? ((wd & 0x0000ffffu) | (ad << 16)) : ((wd & 0xffff0000u) | ad);
    }
  }
  so.push_back(-1000000);
  *ngroup_out = ngroup;
}

namespace {
void read_ffn(const std::string &dir, int iweak0, Network &net) {
  const int nf = std::max(net.nffn, 1);
  net.ffnsum.assign((size_t)nf * kNGrid, 0.0); net.ffnenu.assign((size_t)nf * kNGrid, 0.0);
  net.ffn_ec.assign((size_t)nf * kNGrid, 0.0); net.ffn_beta.assign((size_t)nf * kNGrid, 0.0);
  net.ffn_qval.assign(nf, 0.0); net.has_logft.assign(nf, 0);
  if (net.nffn <= 0) return;
  std::vector<std::string> L = read_lines(dir + "/netweak");
  size_t pos = 0;
  const bool logft = std::abs(iweak0) == 2;
  const int per = logft ? 4 : 2;
  for (int i = 0; i < net.nffn; ++i) {
    if (pos >= L.size()) throw std::runtime_error("netweak: unexpected end of file");
    if (logft) {
      std::istringstream is(L[pos]);
      std::string n1, n2, d;
      std::string qv;
      is >> n1 >> n2 >> d >> qv;
      net.ffn_qval[i] = fortran_real(qv);
      std::string d3 = d.substr(0, 3);
      net.has_logft[i] = d3 == "ft+" ? 2 : (d3 == "ft-" ? 1 : 0);
    }
    ++pos;
    std::vector<double> v;
    while ((int)v.size() < per * kNGrid) {
      if (pos >= L.size()) throw std::runtime_error("netweak: unexpected end of file");
      std::string s = L[pos++];
      for (auto &c : s)
        if (c == ',') c = ' ';
      std::istringstream is(s);
      std::string t;
      while (is >> t) v.push_back(fortran_real(t));
    }
    for (int j = 0; j < kNGrid; ++j) {
      if (logft) {
        net.ffn_beta[(size_t)i * kNGrid + j] = v[4 * j];
        double ec = v[4 * j + 1];
        net.ffn_ec[(size_t)i * kNGrid + j] = ec < -30.0 ? 99.0 : ec;
        net.ffnsum[(size_t)i * kNGrid + j] = v[4 * j + 2];
        net.ffnenu[(size_t)i * kNGrid + j] = v[4 * j + 3];
      } else {
        net.ffnsum[(size_t)i * kNGrid + j] = v[2 * j];
        net.ffnenu[(size_t)i * kNGrid + j] = v[2 * j + 1];
      }
    }
  }
}

}  // namespace

int index_from_name(const Network &net, const std::string &name) {
  // 0 = blank, ny+1 = not in the network (xnet_data.F90:54-82)
  const std::string s = trimmed(name);
  if (s.empty()) return 0;
  for (int i = 0; i < net.ny; ++i)
    if (trimmed(net.nname[i]) == s) return i + 1;
  return net.ny + 1;
}

void load_network(const std::string &data_dir, int iweak0, Network &net) {
  read_nuclear(data_dir, net);
  preprocess_reactions(data_dir, net);
  read_ffn(data_dir, iweak0, net);
  match_react(net);
  structured_lu_plan(net);
}

void read_thermo_file(const std::string &path, Thermo &th) {
  std::vector<std::string> L = read_lines(path);
  if (L.size() < 5) throw std::runtime_error("Failed while trying to read input file: " + path);
  th.desc = field(L[0], 0, 80);
  auto first_tok = [](const std::string &s) { std::istringstream is(s); std::string t; is >> t; return t; };
  th.tstart = fortran_real(first_tok(L[1]));
  th.tstop = fortran_real(first_tok(L[2]));
  th.tdelstart = std::min(0.0, fortran_real(first_tok(L[3])));  // file value is ignored unless negative
  th.th.clear(); th.t9h.clear(); th.rhoh.clear(); th.yeh.clear();
  size_t last = L.size();
  while (last > 4 && L[last - 1].empty()) --last;
  for (size_t i = 4; i < last && th.th.size() < 50000; ++i) {
    std::string s = L[i];
    for (auto &c : s)
      if (c == '\t') c = ' ';
    std::istringstream is(s);
    std::vector<std::string> t;
    std::string x;
    while (is >> x) t.push_back(x);
    if (t.size() < 3) throw std::runtime_error("Not enough columns in thermo file: " + path);
    th.th.push_back(fortran_real(t[0])); th.t9h.push_back(fortran_real(t[1])); th.rhoh.push_back(fortran_real(t[2]));
    th.yeh.push_back(t.size() > 3 ? fortran_real(t[3]) : 0.0);
  }
}

void read_inab_file(const Network &net, const std::string &path, InAb &ab) {
  std::vector<std::string> L = read_lines(path);
  if (L.empty()) throw std::runtime_error("Failed to open input file: " + path);
  ab.desc = field(L[0], 0, 80);
  std::vector<std::string> tok = tokens_from(L, 1);
  const int ny = net.ny;
  ab.y.assign(ny, 0.0);
  ab.yein = 0.0;
  double yext = 0.0;
  for (size_t i = 0; i + 1 < tok.size(); i += 2) {
    std::string nm = tok[i];
    nm.erase(std::remove(nm.begin(), nm.end(), '\''), nm.end());
    nm = trimmed(nm.substr(0, 5));
    const double v = fortran_real(tok[i + 1]);
    if (nm.empty() || !(v > 0.0)) continue;
    for (auto &c : nm)
      if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    if (nm == "ye") { ab.yein = v; continue; }
    const int inuc = index_from_name(net, nm);
    if (inuc < 1 || inuc > ny) yext += v; else ab.y[inuc - 1] = v;
  }
  // sum(yin*aa), sum(yin*zz): gfortran evaluates the array sums left to right
  double xnet = 0.0, znet = 0.0;
  for (int i = 0; i < ny; ++i) { xnet += ab.y[i] * net.aa[i]; znet += ab.y[i] * net.zz[i]; }
  if (yext > 0.0 && xnet < 1.0 && znet < ab.yein) {
    ab.xext = 1.0 - xnet;
    ab.aext = ab.xext / yext;
    ab.zext = (ab.yein <= 0.0 || ab.yein >= 1.0) ? 0.0 : (ab.yein - znet) * ab.aext / ab.xext;
  } else {
    ab.xext = 0.0; ab.aext = 1.0; ab.zext = 0.0;
    for (int i = 0; i < ny; ++i) ab.y[i] = ab.y[i] / xnet;
  }
}

}  // namespace xb

// ---- Helmholtz free-energy table (actual_eos.F90:1609-1652 in the XNet tree) -----------------------------
// ASCII form: 541*201 records of 9 numbers (f fd ft fdd ftt fdt fddt fdtt fddtt), then three blocks of
// 541*201 records of 4 numbers (dpdf*, ef*, xf*); temperature index outermost, density index fastest.
// Only f* and ef* are used by XNet's pruned EOS.  Packed form: those 13 planes as raw little-endian doubles.
namespace xb {
void load_helm_table(const std::string &path, Network &net) {
  const size_t n = (size_t)541 * 201;
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open Helmholtz table " + path);
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<double> planes(13 * n);
  if ((size_t)size == 13 * n * sizeof(double)) {
    const size_t got = std::fread(planes.data(), sizeof(double), 13 * n, f);
    std::fclose(f);
    if (got != 13 * n) throw std::runtime_error("short read on " + path);
  } else {
    std::vector<char> buf((size_t)size + 1);
    const size_t got = std::fread(buf.data(), 1, (size_t)size, f);
    std::fclose(f);
    buf[got] = 0;
    char *p = buf.data();
    auto next = [&](double &v) {
      char *e = nullptr;
      v = std::strtod(p, &e);
      if (e == p) throw std::runtime_error("malformed Helmholtz table " + path);
      p = e;
    };
    double v;
    for (size_t r = 0; r < n; ++r) for (int q = 0; q < 9; ++q) { next(v); planes[(size_t)q * n + r] = v; }
    for (size_t r = 0; r < n; ++r) for (int q = 0; q < 4; ++q) next(v);                 // dpdf block: unused
    for (size_t r = 0; r < n; ++r) for (int q = 0; q < 4; ++q) { next(v); planes[(size_t)(9 + q) * n + r] = v; }
  }
  net.helm.swap(planes);
}
}  // namespace xb
