// Host-side network model of xnet_b200 (C++17).  Mirrors the Fortran modules nuclear_data and
// reaction_data (source/xnet_data.F90:8-52, 602-650 in the XNet tree) as plain vectors.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace xb {

constexpr int kNG = 24;       // partition-function grid points (xnet_data.F90: ng)
constexpr int kNGrid = 143;   // FFN grid 13 x 11 (xnet_ffn.F90:15-17)
constexpr int kNChap = 11;

struct Network {
  int ny = 0, nffn = 0, nnnu = 0, izmax = 0;
  int nreac[4] = {0, 0, 0, 0};
  int nan[4] = {0, 0, 0, 0};
  std::vector<std::string> nname;             // [ny] 5-char names
  std::vector<double> aa, zz, nn, mex, be;    // [ny]
  std::vector<int32_t> iz, in;
  double mex_n = 0, mex_p = 0;
  std::vector<double> g;                      // [ny][24]
  std::vector<double> t9i;                    // [24]
  std::vector<int32_t> chap_la, chap_le;      // [11]
  // per reactant-count class c = 0..3
  std::vector<int32_t> ni[4];                 // [nreac][6]
  std::vector<double> rc[4];                  // [nreac][7]
  std::vector<double> q[4];
  std::vector<int32_t> iwk[4], ires[4], irev[4];
  std::vector<int32_t> iffn;                  // [nreac1]
  std::vector<int32_t> la, le;                // [4][ny] class-major
  std::vector<int32_t> mu[4];                 // [nan]
  std::vector<double> a[4];                   // [nan]
  // reaction-pair ("flux") matching of match_react (xnet_preprocess.F90:108-289) and flux_init (xnet_flux.F90:20-65)
  int mflx = 0;
  std::vector<int32_t> ifl[4];                // [nreac] signed 1-based flux index: + forward (Q > 0), - reverse
  std::vector<int32_t> nflx;                  // [mflx][8] reactants (slots 0-3) and products (4-7), 1-based, 0 = blank
  std::vector<int32_t> iwflx;                 // [mflx] weak flag of the pair
  std::vector<double> qflx;                   // [mflx] |Q|
  std::vector<double> dcflx;                  // [mflx][2] double-counting factors (forward, reverse)
  std::vector<int32_t> ifl_orig, ifl_term;    // [mflx] originating / terminating species (1-based)
  // structured-LU plan (groundwork, not yet used by a kernel): species order with the light particles
  // (Z <= 2, A <= 4) last, and the skyline envelope of the Jacobian's structural pattern in that order, which an LU
  // without interchanges never leaves
  std::vector<int32_t> slu_perm;              // [ny] new position -> original species (1-based)
  std::vector<int32_t> slu_rowfirst;          // [ny] first structural column of row i (1-based, permuted order, <= i)
  std::vector<int32_t> slu_colfirst;          // [ny] first structural row of column j (1-based, permuted order, <= j)
  int64_t slu_doubles = 0;                    // skyline storage: sum (i - rowfirst) + (j - colfirst) + ny
  // FFN tables [nffn][143]
  std::vector<double> ffnsum, ffnenu, ffn_ec, ffn_beta, ffn_qval;
  std::vector<int32_t> has_logft;
  // Helmholtz table planes [13][201][541]: f fd ft fdd ftt fdt fddt fdtt fddtt ef efd eft efdt (empty = not loaded)
  std::vector<double> helm;
};

// throws std::runtime_error on malformed input
void load_network(const std::string &data_dir, int iweak0, Network &net);
// fills mflx, ifl, nflx, iwflx, qflx, dcflx, ifl_orig, ifl_term from chap_la/chap_le, ni, q, iwk
void match_react(Network &net);
// dcflx, ifl_orig, ifl_term from mflx/nflx (second half of match_react; also used for networks built from arrays)
void flux_init(Network &net);
// fills slu_perm, slu_rowfirst, slu_colfirst, slu_doubles from the species <-> reaction maps
void structured_lu_plan(Network &net);
// Tables of the structured LU for a system of N = ny (+1 with self-heating: the T9 equation is the last row and
// column, both full) equations in the light-particles-last order: nh = n - nbord "heavy" equations with a variable-band
// pattern, then the border (light particles, T9).  Storage, S[nstore] doubles:
//   rows i < nh  : COLUMN-major, column j holds rows ct[j] = cf[j] .. cbot[j] contiguously, element (i,j) at cb[j] + i.
//                  Heavy columns: cbot[j] = last heavy row whose stored part reaches column j (>= j); border columns: all
//                  heavy rows.  Column starts are padded so that the bases cb[j] of any 8 (XB_SLU_PADW) consecutive columns differ mod 16:
//                  lanes working on consecutive columns at one row hit distinct shared-memory banks.
//   rows i >= nh : also by column, element (i,j) at bo + j * bs + (i - nh) with an odd stride bs >= nbord; n + 1 columns, the
//                  last one carries the border rows' right-hand side entries through the factorisation.
// Elimination without interchanges over the heavy part never leaves this envelope (it contains the row/column skyline
// rf/cf of structured_lu_plan; the zero entries it adds stay zero).
struct SluTables {
  int n = 0, nstore = 0, nbord = 0;           // order, doubles of the store, trailing rows/columns factored with pivoting
  int bo = 0, bs = 0;                         // border rows: base and column stride
  std::vector<int32_t> perm, pos;             // [n] position -> equation, equation -> position (0-based; equation ny = T9)
  std::vector<int32_t> rf, cf;                // [n] first structural column of row i / first structural row of column j
  std::vector<int32_t> cb, cbot;              // [n] column base (address of a virtual row 0), last stored heavy row
  std::vector<int32_t> nrb, ncb;              // [nh] per pivot k: rows k+1..k+nrb and columns k+1..k+ncb of the heavy part take part
  int pair = 0;                               // XB_SLU_PAIR=1: profiles rounded down to even, pivots (k, k+1), k even, share rows and columns (rank-2 rounds; measured slower, off)
  double flops = 0;                           // of one factorisation over the structural pattern (2 per multiply-add, 1 per multiplier)
  double flops_envelope = 0;                  // what the kernel executes (spans instead of patterns)
  bool inside(int i, int j) const {
    const int nh = n - nbord;
    return i >= nh ? true : (cf[j] <= i && i <= cbot[j]);
  }
  int addr(int i, int j) const {
    const int nh = n - nbord;
    return i >= nh ? bo + j * bs + (i - nh) : cb[j] + i;
  }
};
void structured_lu_tables(const Network &net, int heat, SluTables &t);
// Pivot schedule of the structured LU: everything slu_factor's index arithmetic would derive per pivot, per warp and per
// lane, laid out as the kernel reads it.  One ROUND = one barrier interval in which each of the nwl factoring warps
// updates one chunk of up to 8 rows (band rows of the pivot, or the border rows) of up to 32 columns (lane = column:
// the first 16 band columns, the border columns and the right-hand side, further band columns in extra rounds).  A
// pivot takes one round unless it has more chunks than warps or more columns than lanes.  All addresses are double
// indices into the store S (the right-hand side = the pseudo-column at rhs_off).
//   warp[(r * nwl + w) * 2]     = lp | pv << 16   lp: multipliers of the chunk (column k), pv: the pivot a(k,k)
//   warp[(r * nwl + w) * 2 + 1] = nrow | ntest << 4 | k << 8   rows of the chunk, how many of them are subject to the
//                                                  threshold test (0 in the rounds that revisit a chunk), the pivot
//   lane[(r * nwl + w) * 32 + l] = ap | up << 16  ap: first row of the chunk in this lane's column, up: u(k, column);
//                                                  ap == lp marks an idle lane (it loads, stores nothing)
// nround is padded to a multiple of three with idle rounds and three more idle rounds follow (the kernel runs three
// rounds per loop trip and prefetches three rounds ahead).
struct SluSchedule {
  int nround = 0, nwl = 0;
  std::vector<uint32_t> warp, lane;
};
// ntest_rows: equations 0..ntest_rows-1 are tested (all but the T9 row)
void structured_lu_schedule(const SluTables &t, int nwl, int rhs_off, int ntest_rows, SluSchedule &s);
// Rank-2 schedule (pair-aligned tables): one round eliminates the pivots (k, k+1), k even (a last odd pivot runs as a
// "single": its second pivot is switched off).  Chunks cover the rows k+2.. (single: k+1..) and the border rows; row k+1,
// i.e. U(k+1, .), is formed by every lane for its own column and stored, one barrier later, by the lanes of the item
// flagged `commit`.  Per warp and round four words:
//   [0] lp1 | lp2 << 16         the chunk's rows in column k and in column k+1
//   [1] nrow | ntest1 << 4 | ntest2 << 8 | commit << 12 | single << 13
//   [2] pv1 | u12 << 16         a(k,k) and a(k,k+1); a(k+1,k) = pv1 + 1, a(k+1,k+1) = u12 + 1
//   [3] k
// per lane and round two words:  [0] ap | up1 << 16  (column segment, a(k,j));  [1] up2 | col2 << 16  (a(k+1,j); col2: j = k+1,
// the lane that leaves the updated multipliers of pivot k+1 behind);  ap == lp1 marks an idle lane.  Padding as above.
struct SluSchedule2 {
  int nround = 0, nwl = 0;
  std::vector<uint32_t> warp, lane;
};
void structured_lu_schedule2(const SluTables &t, int nwl, int rhs_off, int ntest_rows, SluSchedule2 &s);
// Band back-substitution table (slu_solve, phase C): per step k = nh-1 .. 0 and lane the address of U(i,k) for the window row
// i = k - ((k - lane) & 31), of the pivot's reciprocal for i = k, 0xffff where the column does not reach; four 16-bit entries
// per lane and group of four steps: lane_words[(g * 32 + lane) * 2 + s / 2], low half first; ngroup is even and two idle
// groups follow.  out_steps: the steps whose column reaches above the window (cf[k] <= k - 32), descending, then -1000000.
void structured_lu_solve_table(const SluTables &t, std::vector<uint32_t> &lane_words, std::vector<int32_t> &out_steps, int *ngroup);
int index_from_name(const Network &net, const std::string &name);
void load_helm_table(const std::string &path, Network &net);

struct Thermo {
  std::string desc;
  double tstart = 0, tstop = 0, tdelstart = 0;
  std::vector<double> th, t9h, rhoh, yeh;
};
void read_thermo_file(const std::string &path, Thermo &th);

struct InAb {
  std::string desc;
  double yein = 0, xext = 0, aext = 1, zext = 0;
  std::vector<double> y;
};
void read_inab_file(const Network &net, const std::string &path, InAb &ab);

}  // namespace xb
