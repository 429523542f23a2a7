// search_or_cuda.cc -- see search_or_cuda.h.  Host glue only: every number comes from the
// CUDA library (no CPU evaluation of the search lives here).
#include "search_or_cuda.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include <distance_metric.h>
#include <global_utils.h>

using namespace std;

namespace minrms {

static void Check(mr_ctx* c, int rc)
{
  if (rc != 0) ERR("CUDA search failed: " << mr_last_error(c));  // ERR.h:20-21 -> exit(-1)
}

SearchOrCuda::SearchOrCuda(SearchOrDyn3d::Settings const& s) : SearchOrDyn3d(s)
{
  if (settings.num_solutions != 1)
    ERR("Error: alternate solutions (-alt-o, -alt-r) are not supported by the CUDA search.");
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  N = max_num_matches;
  R = range_of_n;
  mr_params p;
  p.nb = nb;
  p.n_min = n_min;
  p.n_max = n_max;
  p.max_dist = settings.dyn3d_max_dist;
  p.max_rmsd = settings.dyn3d_max_rmsd;
  p.use_filter = settings.orientation_filter_nw ? 1 : 0;
  p.of_dist = settings.orientation_filter_nw_max_dist;
  p.of_min = settings.orientation_filter_nw_min_num_matches;
  p.refine_while_searching = settings.refine_while_searching ? 1 : 0;
  int ngpu = 1;
  if (const char* e = getenv("MINRMS_CUDA_GPUS")) ngpu = max(1, atoi(e));
  // the flattened coordinate arrays the reference builds at search_or_dyn3d.cc:186-224
  for (int g = 0; g < ngpu; ++g) {
    mr_ctx* c = NULL;
    if (mr_create(&c, &aRotateUsingTheseCoords1[0][0], numRes1, &aRotateUsingTheseCoords2[0][0],
                  numRes2, &p, g) != 0)
      ERR("Error: cannot start the CUDA search: " << mr_last_error(NULL));
    ctx.push_back(c);
  }
  ssd.resize(R);
  T.resize(12 * R);
  pairs.resize((size_t)2 * N * R);
  oid.resize(R);
}

SearchOrCuda::~SearchOrCuda()
{
  for (size_t g = 0; g < ctx.size(); ++g) mr_destroy(ctx[g]);
}

void SearchOrCuda::TableToHost(mr_best& t)
{
  t.ssd = &ssd[0]; t.T = &T[0]; t.pairs = &pairs[0]; t.o_id = &oid[0];
  for (int r = 0; r < R; ++r) {
    ssd[r] = aaSumSqdDists[0][r];
    memcpy(&T[12 * r], &aaTransforms[0][r][0][0], 12 * sizeof(float));
    const int n = n_min + r;
    for (int k = 0; k < n; ++k) {
      pairs[(size_t)2 * N * r + 2 * k] = aaAlignments[0][r][k][0];
      pairs[(size_t)2 * N * r + 2 * k + 1] = aaAlignments[0][r][k][1];
    }
  }
}

void SearchOrCuda::TableFromHost(mr_best const& t)
{
  for (int r = 0; r < R; ++r) {
    aaSumSqdDists[0][r] = t.ssd[r];
    memcpy(&aaTransforms[0][r][0][0], &t.T[12 * r], 12 * sizeof(float));
    const int n = n_min + r;
    if (!(t.ssd[r] < DistanceMetric::UPPER_BOUND)) continue;
    for (int k = 0; k < n; ++k) {
      aaAlignments[0][r][k][0] = t.pairs[(size_t)2 * N * r + 2 * k];
      aaAlignments[0][r][k][1] = t.pairs[(size_t)2 * N * r + 2 * k + 1];
    }
  }
}

// FirstPass over og[start, stop) (search_or_dyn3d.cc:1643-1946).  vector<OrGenEntry> is a
// contiguous array of 12-float Matrix3x4 (or_gen.h:73-78), so the range is passed as is;
// with several GPUs it is cut into contiguous slices, one host thread per GPU, and the
// per-size tables are merged with the reference's "lower score, then lower id" rule.
void SearchOrCuda::Search(OrientationGenerator const& og, long start, long stop)
{
  const int G = (int)ctx.size();
  vector<vector<float> > vs(G), vT(G);
  vector<vector<int16_t> > vp(G);
  vector<vector<int64_t> > vo(G);
  vector<mr_best> tb(G);
  vector<int> rc(G, 0);
  vector<thread> th;
  for (int g = 0; g < G; ++g) {
    vs[g].resize(R); vT[g].resize(12 * R); vp[g].assign((size_t)2 * N * R, 0); vo[g].resize(R);
    tb[g].ssd = &vs[g][0]; tb[g].T = &vT[g][0]; tb[g].pairs = &vp[g][0]; tb[g].o_id = &vo[g][0];
    mr_best_init(ctx[g], &tb[g]);
    const long a = start + (stop - start) * g / G, b = start + (stop - start) * (g + 1) / G;
    const float* first = (b > a) ? &og[a].transform[0][0] : NULL;
    th.push_back(thread([=, &rc, &tb]() { rc[g] = (b > a) ? mr_search(ctx[g], first, b - a, a, &tb[g]) : 0; }));
  }
  for (int g = 0; g < G; ++g) th[g].join();
  for (int g = 0; g < G; ++g) Check(ctx[g], rc[g]);
  mr_best mine;
  TableToHost(mine);
  for (int g = 0; g < G; ++g)
    for (int r = 0; r < R; ++r) {
      const bool better = tb[g].ssd[r] < mine.ssd[r] ||
                          (tb[g].ssd[r] == mine.ssd[r] && tb[g].ssd[r] < DistanceMetric::UPPER_BOUND &&
                           oid[r] >= 0 && tb[g].o_id[r] < oid[r]);
      if (!better) continue;
      mine.ssd[r] = tb[g].ssd[r];
      oid[r] = tb[g].o_id[r];
      memcpy(&mine.T[12 * r], &tb[g].T[12 * r], 12 * sizeof(float));
      memcpy(&mine.pairs[(size_t)2 * N * r], &tb[g].pairs[(size_t)2 * N * r], sizeof(int16_t) * 2 * (n_min + r));
    }
  TableFromHost(mine);
  int n;
  for (n = n_min; n <= n_max && aaSumSqdDists[0][n - n_min] == DistanceMetric::UPPER_BOUND; ++n) {}
  if (n > n_max)  // search_or_dyn3d.cc:1929-1941
    ERR("Error: No solutions were found satisfying your criteria.\n"
        "       The restrictions you placed on accepting solutions were\n"
        "       too strict.  Try using a lower minimum-number-of-matches,\n"
        "       a larger maximum distance cutoff, for example.");
}

// search_or_dyn3d.cc:2261-2658 (REFINE_BEST)
void SearchOrCuda::Refine()
{
  cout << "Entering refinement stage." << endl;
  mr_best mine;
  TableToHost(mine);
  for (int r = 0; r < R; ++r) oid[r] = (mine.ssd[r] < DistanceMetric::UPPER_BOUND) ? 0 : -1;
  int32_t evals = 0;
  const long iters = settings.refine_max_iters;  // ITERATE_UNTIL_CONVERGENCE = -1
  Check(ctx[0], mr_refine(ctx[0], &mine, (int32_t)iters, &evals));
  TableFromHost(mine);
  cout << "(During refinement, " << evals + 1 << " new orientations were tested on the GPU.)" << endl;
}

// search_or_dyn3d.cc:2942-2996.  The rows of the plot (WritePlotFile, 2662-2760: longest
// matched distance, Levitt-Gerstein probability) come from mr_plot_stats; the streaming and
// number formatting are the reference's (default ostream precision, same value types).
void SearchOrCuda::WriteChimeraInfo() const
{
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  int ca = -1;
  for (int q = 0; q < nb; ++q)
    if (strcmp(Biopolymer::Residue::LookupBackboneAtomSymbol(q), " CA") == 0) ca = q;
  if (ca < 0) {  // S_str needs C-alpha atoms that the search never moved: the reference's own path
    SearchOrDyn3d::WriteChimeraInfo();
    return;
  }
  WritePlotFileCuda(ca);
  string info_name = string(SearchOrDyn3d::BASE_FILE_NAME) + "_chimera.info";
  ofstream info_file(info_name.c_str(), ios::out);
  if (!info_file) ERR_INTERNAL("Cannot open file \"" << info_name << "\" for writing.  Exiting...");
  info_file << "fixed pdb:" << settings.vPdbFileNames[0] << "\n";
  info_file << "moveable pdb:" << settings.vPdbFileNames[1] << "\n";
  info_file << "rmsd:" << SearchOrDyn3d::BASE_FILE_NAME << "_chimera.plot\n";
  info_file << "base:" << SearchOrDyn3d::BASE_FILE_NAME << "\n";
  info_file << "number of atoms used:" << nb << "\n";
  for (int q = 0; q < nb; ++q) info_file << Biopolymer::Residue::LookupBackboneAtomSymbol(q) << "\n";
}

void SearchOrCuda::WritePlotFileCuda(int ca_index) const
{
  vector<float> vs(R), vT(12 * R);
  vector<int16_t> vp((size_t)2 * N * R, 0);
  for (int r = 0; r < R; ++r) {
    vs[r] = aaSumSqdDists[0][r];
    memcpy(&vT[12 * r], &aaTransforms[0][r][0][0], 12 * sizeof(float));
    if (!(vs[r] < DistanceMetric::UPPER_BOUND)) continue;
    for (int k = 0; k < n_min + r; ++k) {
      vp[(size_t)2 * N * r + 2 * k] = aaAlignments[0][r][k][0];
      vp[(size_t)2 * N * r + 2 * k + 1] = aaAlignments[0][r][k][1];
    }
  }
  mr_best t;
  t.ssd = &vs[0]; t.T = &vT[0]; t.pairs = &vp[0]; t.o_id = NULL;
  vector<double> rows((size_t)9 * R);
  Check(ctx[0], mr_plot_stats(ctx[0], &t, ca_index, &rows[0]));
  string plot_name = string(SearchOrDyn3d::BASE_FILE_NAME) + "_chimera.plot";
  ofstream plot_file(plot_name.c_str(), ios::out);
  if (!plot_file) ERR_INTERNAL("WritePlotFile() cannot create file \"" << plot_name << "\" for writing.");
  plot_file << "# Matches,RMSD,Longest Distance" << ",-log(probability)\n";
  for (int r = 0; r < R; ++r) {
    const double* o = &rows[(size_t)9 * r];
    if (o[0] == 0.0) continue;
    // rmsd and distance are Real in the reference, the last column a double (2719-2755)
    plot_file << n_min + r << " " << (float)o[1] << " " << (float)o[2] << " " << o[8] << "\n";
  }
}

}  // namespace minrms
