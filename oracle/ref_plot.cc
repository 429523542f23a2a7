// ref_plot.cc -- TEST INFRASTRUCTURE.  Built only where /root/reference exists
// (oracle/build_ref.sh), linked against the UNMODIFIED reference objects.
//
// Usage: ref_plot IN.bin OUT.bin <minrms arguments...>
//
// Evaluates, with the reference's own code, the per-size statistics that
// SearchOrDyn3d::WritePlotFile prints into align_chimera.plot
// (bin/minrms/search_or_dyn3d.cc:2703-2760) for a given best table:
//   IN.bin : int32 R, int32 N, then R records {float ssd, float T[12], int16 pairs[N][2]}
//            (slot r = size n_min + r, the layout of mr_best)
//   OUT.bin: R records of 9 doubles {has_solution, rmsd as printed, furthest distance,
//            S_str, mu_str, delta_str, Z, probability, -log10(probability) as printed}
//            followed by 2 int32: sizeof() of the printed rmsd and -log10 expressions.
// All numbers are produced by reference code; this file is only glue + I/O.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <iostream>
using namespace std;

#include <global_utils.h>
#include <parse_utils.h>
#include <vect_3d.h>
#include <biopolymer.h>
#include <apply_trans.h>
#include <pair_alignment.h>
#include <distance_metric.h>
#include "search_or.h"
#include "search_or_dyn3d.h"
#include "minrms_parser.h"

using namespace minrms;

extern const char g_version_string[] = "ref_plot";
extern const char g_date_string[] = "<oracle>";

int main(int argc, char **argv)
{
  if (argc < 4) { fprintf(stderr, "usage: ref_plot IN.bin OUT.bin <minrms args>\n"); return 2; }
  vector<string> vOrig;
  vOrig.push_back("minrms");
  for (int i = 3; i < argc; i++) vOrig.push_back(argv[i]);
  char back_slash[2] = {'\\', '\0'};
  vector<string> vArgs;
  ParseUtils::HandleArgFiles(vOrig, vArgs, "-cfg", "#", back_slash, string());
  MinrmsParser settings(vArgs);
  if (vArgs.size() != 1) { fprintf(stderr, "unrecognized option %s\n", vArgs[1].c_str()); return 2; }
  const int n_min = settings.n_min;

  FILE *in = fopen(argv[1], "rb");
  if (!in) { perror(argv[1]); return 2; }
  int R = 0, N = 0;
  if (fread(&R, 4, 1, in) != 1 || fread(&N, 4, 1, in) != 1) return 2;
  FILE *out = fopen(argv[2], "wb");

  Biopolymer mol_c[2];
  mol_c[0] = settings.aMol_f[0];
  mol_c[1] = settings.aMol_f[1];
  int size_rmsd = 0, size_nl = 0;
  for (int r = 0; r < R; r++) {
    const int n = n_min + r;
    float ssd;
    Matrix3x4 T;
    vector<short> pairs(2 * N);
    if (fread(&ssd, 4, 1, in) != 1 || fread(&T[0][0], 4, 12, in) != 12 ||
        fread(&pairs[0], 2, 2 * N, in) != (size_t)(2 * N)) return 2;
    double rec[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (ssd < DistanceMetric::UPPER_BOUND) {
      PairwiseAlignment al(n);
      al.resize(n);
      for (int k = 0; k < n; k++) { al[k][0] = pairs[2 * k]; al[k][1] = pairs[2 * k + 1]; }
      rec[0] = 1.0;
      // the expressions of WritePlotFile, verbatim
      rec[1] = sqrt(ssd / (n * Biopolymer::Residue::NumBackboneAtoms()));
      size_rmsd = sizeof(sqrt(ssd / (n * Biopolymer::Residue::NumBackboneAtoms())));
      ApplyTransformBackbone(T, settings.aMol_f[1], mol_c[1]);
      Biopolymer::const_iterator worst_i, worst_j;
      rec[2] = al.FindFurthestResPair(worst_i, worst_j, mol_c[0], mol_c[1]);
      Real S_str, mu_str, delta_str, Z;
      Real probability = LevittGerstein98::Sstr_Probability(al, mol_c[0], mol_c[1], &S_str, &mu_str, &delta_str, &Z);
      rec[3] = S_str; rec[4] = mu_str; rec[5] = delta_str; rec[6] = Z; rec[7] = probability;
      rec[8] = ((probability == 0.0) ? 0.0 : -log10(probability));
      size_nl = sizeof(((probability == 0.0) ? 0.0 : -log10(probability)));
    }
    fwrite(rec, 8, 9, out);
  }
  fwrite(&size_rmsd, 4, 1, out);
  fwrite(&size_nl, 4, 1, out);
  fclose(out);
  fclose(in);
  return 0;
}
