// ref_dump.cc -- TEST INFRASTRUCTURE.  Built only where /root/reference exists
// (oracle/build_ref.sh), linked against the UNMODIFIED reference objects.
//
// Usage: ref_dump OUT.bin DETAIL_STRIDE <minrms arguments...>
//
// Parses the command line with the reference's own MinrmsParser, builds the
// reference's OrientationGenerator, then
//   (1) drives the reference's own Dyn3d / OrientationFilter / PairwiseAlignment /
//       Superimpose objects in the order FirstPass does
//       (bin/minrms/search_or_dyn3d.cc:1691-1835) and dumps, per orientation, the
//       values the restatement in oracle/minrms_oracle.c must reproduce;
//   (2) runs the reference's real SearchOrDyn3d::Search() and (if enabled) Refine()
//       and dumps the best tables after each.
// All numbers are produced by reference code; this file is only glue + I/O.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <iostream>
using namespace std;

#define private public   // read SearchOrDyn3d's result tables (test-only hack)
#include <global_utils.h>
#include <parse_utils.h>
#include <or_gen.h>
#include <orientation_filter.h>
#include <dyn3d.h>
#include <distance_metric.h>
#include <simple_numeric_utils.h>
#include <apply_trans.h>
#include <superimpose.h>
#include "search_or.h"
#include "search_or_dyn3d.h"
#include "minrms_parser.h"
#undef private

using namespace minrms;

// main.cc defines these for the writers (search_or_dyn3d.cc); provide them here.
extern const char g_version_string[] = "ref_dump";
extern const char g_date_string[] = "<oracle>";

static FILE *g_out;
static void w_i32(int v) { fwrite(&v, 4, 1, g_out); }
static void w_i64(long long v) { fwrite(&v, 8, 1, g_out); }
static void w_f32(float v) { fwrite(&v, 4, 1, g_out); }
static void w_mat(ConstMatrix3x4 M) { for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) w_f32(M[r][c]); }

static void dump_best(SearchOrDyn3d &s, int n_min, int n_max, int N)
{
  for (int n = n_min; n <= n_max; n++) {
    w_f32(s.aaSumSqdDists[0][n - n_min]);
    w_mat(s.aaTransforms[0][n - n_min]);
    PairwiseAlignment const &al = s.aaAlignments[0][n - n_min];
    for (int k = 0; k < N; k++) {
      short i = -1, j = -1;
      if (k < n && s.aaSumSqdDists[0][n - n_min] < DistanceMetric::UPPER_BOUND) { i = al[k][0]; j = al[k][1]; }
      fwrite(&i, 2, 1, g_out);
      fwrite(&j, 2, 1, g_out);
    }
  }
}

int main(int argc, char **argv)
{
  if (argc < 4) { fprintf(stderr, "usage: ref_dump OUT.bin DETAIL_STRIDE <minrms args>\n"); return 2; }
  const char *out_name = argv[1];
  int stride = atoi(argv[2]);
  int skip_per_orientation = (stride < 0);
  vector<string> vOrig;
  vOrig.push_back("minrms");
  for (int i = 3; i < argc; i++) vOrig.push_back(argv[i]);
  char back_slash[2] = {'\\', '\0'};
  vector<string> vArgs;
  ParseUtils::HandleArgFiles(vOrig, vArgs, "-cfg", "#", back_slash, string());
  MinrmsParser settings(vArgs);
  if (vArgs.size() != 1) { fprintf(stderr, "unrecognized option %s\n", vArgs[1].c_str()); return 2; }

  OrientationGenerator og(settings);
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  const int m = settings.aMol_f[0].size(), nB = settings.aMol_f[1].size();
  const int N = (m < nB) ? m : nB;
  const int n_min = settings.n_min, n_max = settings.n_max;
  const long count = og.size();

  g_out = fopen(out_name, "wb");
  w_i32(0x4d524d53); // magic
  w_i32(m); w_i32(nB); w_i32(nb); w_i32(n_min); w_i32(n_max);
  w_f32(settings.dyn3d_max_dist); w_f32(settings.dyn3d_max_rmsd);
  w_i32(settings.orientation_filter_nw ? 1 : 0);
  w_f32(settings.orientation_filter_nw_max_dist);
  w_i32(settings.orientation_filter_nw_min_num_matches);
  w_i32(settings.refine_while_searching ? 1 : 0);
  w_i32(settings.refine_method != SearchOr::Settings::NO_REFINE ? 1 : 0);
  w_i32(settings.fm_fragment_size);
  w_i64(count);
  w_i32(stride);

  // coordinates as the hot path sees them (search_or_dyn3d.cc:186-224)
  for (int mol = 0; mol < 2; mol++)
    for (Biopolymer::const_iterator ps = settings.aMol_f[mol].begin(); ps < settings.aMol_f[mol].end(); ++ps)
      for (int q = 0; q < nb; q++) {
        Biopolymer::Residue::const_iterator pa = ps->GetBackboneAtom(q);
        w_f32((*pa).second.xyz[0]); w_f32((*pa).second.xyz[1]); w_f32((*pa).second.xyz[2]);
      }
  for (long o = 0; o < count; o++) w_mat(og[o].transform);

  if (!skip_per_orientation) {
    // ---- (1) per-orientation values from the reference's own components ----
    Dyn3d dyn3d(m, nB, n_min, n_max);
    dyn3d.pCostFunc = &DistanceMetric::SquaredDistanceBetweenIandJ;
    DistanceMetric::SetMaxPhysicalDistance(settings.dyn3d_max_dist);
    Biopolymer mol_c[2];
    mol_c[0] = settings.aMol_f[0];
    mol_c[1] = settings.aMol_f[1];
    Superimpose sup(N * nb);
    OrientationFilter *pFilter = NULL;
    if (settings.orientation_filter_nw)
      pFilter = new OrientationFilter(settings.orientation_filter_nw_max_dist,
                                      settings.orientation_filter_nw_min_num_matches, m, nB);
    Vect3 *g1 = new Vect3[m * nb], *g2 = new Vect3[nB * nb];
    PairwiseAlignment al(N);
    vector<float> cost(N), ssd(N), T(12 * N);
    vector<unsigned char> valid(N);
    vector<short> pairs(2 * N * N);
    for (long o = 0; o < count; o++) {
      ApplyTransformBackbone(og[o].transform, settings.aMol_f[1], mol_c[1]);
      int good = 1;
      if (pFilter) good = pFilter->GoodOrientation(mol_c[0], mol_c[1]) ? 1 : 0;
      int n_last = 0;
      std::fill(valid.begin(), valid.end(), 0);
      std::fill(cost.begin(), cost.end(), 0.0f);
      std::fill(ssd.begin(), ssd.end(), 0.0f);
      if (good) {
        dyn3d.FillLookupTable(mol_c[0], mol_c[1]);
        dyn3d.InitTable_IncrN();
        cost[0] = dyn3d.aOutcomes[0].cost;
        n_last = 1;
        for (int n = 2; n <= n_max; n++) {
          dyn3d.Solve_IncrN(n);
          cost[n - 1] = dyn3d.aOutcomes[n - 1].cost;
          n_last = n;
          if (dyn3d.MaxDistanceExceeded(n)) break;
          if (n >= n_min) {
            al.resize(n);
            dyn3d.ExportSolution(n, al);
            Matrix3x4 Tn;
            float s;
            if (settings.refine_while_searching) {
              Real rmsd = al.CalcMinRMSD(settings.aMol_f[0], settings.aMol_f[1], Tn, &sup, g1, g2);
              s = SQR(rmsd) * n * nb;
            } else {
              s = dyn3d.GetCost(n);
              CopyMatrix3x4(og[o].transform, Tn);
            }
            if ((settings.dyn3d_max_rmsd != settings.dyn3d_NO_MAX_RMSD) &&
                (s > n * nb * SQR(settings.dyn3d_max_rmsd)))
              break;
            ssd[n - 1] = s;
            valid[n - 1] = 1;
            for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T[12 * (n - 1) + 4 * r + c] = Tn[r][c];
            for (int k = 0; k < n; k++) { pairs[(n - 1) * 2 * N + 2 * k] = al[k][0]; pairs[(n - 1) * 2 * N + 2 * k + 1] = al[k][1]; }
          }
        }
      }
      w_i32(good);
      w_i32(n_last);
      fwrite(&cost[0], 4, N, g_out);
      fwrite(&valid[0], 1, N, g_out);
      fwrite(&ssd[0], 4, N, g_out);
      if (stride > 0 && o % stride == 0) {
        fwrite(&T[0], 4, 12 * N, g_out);
        fwrite(&pairs[0], 2, 2 * N * N, g_out);
      }
    }
  }

  // ---- (2) the real search ----
  SearchOrDyn3d search(settings);
  search.Search(og, 0, og.size());
  dump_best(search, n_min, n_max, N);
  if (settings.refine_method != SearchOr::Settings::NO_REFINE) {
    search.Refine();
    dump_best(search, n_min, n_max, N);
  }
  fclose(g_out);
  return 0;
}
