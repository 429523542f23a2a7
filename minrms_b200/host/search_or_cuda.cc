// search_or_cuda.cc -- see search_or_cuda.h.  Host glue only: every number comes from the
// CUDA library (no CPU evaluation of the search lives here).
#include "search_or_cuda.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

#include <distance_metric.h>
#include <global_utils.h>
#include <simple_numeric_utils.h>  // SQR

using namespace std;

namespace minrms {

static void Check(mr_ctx* c, int rc)
{
  if (rc != 0) ERR("CUDA search failed: " << mr_last_error(c));  // ERR.h:20-21 -> exit(-1)
}

// Score columns of sizes row0 .. row0 + cols.size() - 1 out of the [count][R] array of
// mr_search_scores: one sequential pass over the array per block of 16 sizes (a cache line of a
// row) instead of a strided walk per size -- the array is 0.86 GB at config-3 size.
static const int kScoreBlock = 16;
static void GatherScoreColumns(vector<float> const& scores, long count, int R, int row0,
                               vector<vector<float> >& cols)
{
  const int nrows = (int)cols.size();
  for (int c = 0; c < nrows; ++c) cols[c].resize(count);
  for (long o = 0; o < count; ++o) {
    const float* src = &scores[(size_t)o * R + row0];
    for (int c = 0; c < nrows; ++c) cols[c][o] = src[c];
  }
}

SearchOrCuda::SearchOrCuda(SearchOrDyn3d::Settings const& s) : SearchOrDyn3d(s)
{
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
  coordsA.assign(&aRotateUsingTheseCoords1[0][0], &aRotateUsingTheseCoords1[0][0] + (size_t)3 * nb * numRes1);
  coordsB.assign(&aRotateUsingTheseCoords2[0][0], &aRotateUsingTheseCoords2[0][0] + (size_t)3 * nb * numRes2);
  for (int g = 0; g < ngpu; ++g) {
    mr_ctx* c = NULL;
    if (mr_create(&c, &aRotateUsingTheseCoords1[0][0], numRes1, &aRotateUsingTheseCoords2[0][0],
                  numRes2, &p, g) != 0)
      ERR("Error: cannot start the CUDA search: " << mr_last_error(NULL));
    // work space of the context (default: sized by the library, up to 32 GB, allocated at the first
    // search): runs that start many short processes on one GPU (batch_cuda.cc) set a smaller one
    if (const char* e = getenv("MINRMS_CUDA_BUDGET_MB"))
      if (atol(e) > 0 && mr_set_option(c, "budget_mb", atol(e)) != 0)
        ERR("Error: cannot set the CUDA work space: " << mr_last_error(c));
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

void SearchOrCuda::TableToHost(mr_best& t, int rank)
{
  t.ssd = &ssd[0]; t.T = &T[0]; t.pairs = &pairs[0]; t.o_id = &oid[0];
  for (int r = 0; r < R; ++r) {
    ssd[r] = aaSumSqdDists[rank][r];
    memcpy(&T[12 * r], &aaTransforms[rank][r][0][0], 12 * sizeof(float));
    const int n = n_min + r;
    if (rank > 0 && !(ssd[r] < DistanceMetric::UPPER_BOUND)) continue;  // alternates without a solution keep an unsized alignment
    for (int k = 0; k < n; ++k) {
      pairs[(size_t)2 * N * r + 2 * k] = aaAlignments[rank][r][k][0];
      pairs[(size_t)2 * N * r + 2 * k + 1] = aaAlignments[rank][r][k][1];
    }
  }
}

void SearchOrCuda::TableFromHost(mr_best const& t, int rank)
{
  for (int r = 0; r < R; ++r) {
    aaSumSqdDists[rank][r] = t.ssd[r];
    const int n = n_min + r;
    if (!(t.ssd[r] < DistanceMetric::UPPER_BOUND)) continue;
    memcpy(&aaTransforms[rank][r][0][0], &t.T[12 * r], 12 * sizeof(float));
    aaAlignments[rank][r].resize(n);
    for (int k = 0; k < n; ++k) {
      aaAlignments[rank][r][k][0] = t.pairs[(size_t)2 * N * r + 2 * k];
      aaAlignments[rank][r][k][1] = t.pairs[(size_t)2 * N * r + 2 * k + 1];
    }
  }
}

// One search over `count` transforms (ids id0, id0 + 1, ...) into the tables of `rank`: with
// several GPUs the list is cut into contiguous slices, one host thread per GPU, and the per-size
// tables are merged with the reference's "lower score, then lower id" rule
// (search_or_dyn3d.cc:1825-1835).  scores: mr_search_scores; n_lower > 0: mr_search_dissimilar.
void SearchOrCuda::SearchRank(int rank, const float* first12, long count, long id0, float* scores,
                              int n_lower, const mr_best* lower, vector<int64_t>& ids)
{
  const int G = (int)ctx.size();
  vector<vector<float> > vs(G), vT(G);
  vector<vector<int16_t> > vp(G);
  vector<vector<int64_t> > vo(G);
  vector<mr_best> tb(G);
  vector<int> rc(G, 0);
  vector<thread> th;
  const int max_common = settings.alt_max_pairs_common;
  for (int g = 0; g < G; ++g) {
    vs[g].resize(R); vT[g].resize(12 * R); vp[g].assign((size_t)2 * N * R, 0); vo[g].resize(R);
    tb[g].ssd = &vs[g][0]; tb[g].T = &vT[g][0]; tb[g].pairs = &vp[g][0]; tb[g].o_id = &vo[g][0];
    mr_best_init(ctx[g], &tb[g]);
    const long a = count * g / G, b = count * (g + 1) / G;
    const float* first = (b > a) ? first12 + 12 * a : NULL;
    float* sc = scores && b > a ? scores + (size_t)a * R : NULL;
    th.push_back(thread([=, &rc, &tb]() {
      if (b <= a) rc[g] = 0;
      else if (n_lower > 0) rc[g] = mr_search_dissimilar(ctx[g], first, b - a, id0 + a, &tb[g], n_lower, lower, max_common);
      else if (sc) rc[g] = mr_search_scores(ctx[g], first, b - a, id0 + a, &tb[g], sc);
      else rc[g] = mr_search(ctx[g], first, b - a, id0 + a, &tb[g]);
    }));
  }
  for (int g = 0; g < G; ++g) th[g].join();
  for (int g = 0; g < G; ++g) Check(ctx[g], rc[g]);
  mr_best mine;
  TableToHost(mine, rank);
  for (int g = 0; g < G; ++g)
    for (int r = 0; r < R; ++r) {
      const bool better = tb[g].ssd[r] < mine.ssd[r] ||
                          (tb[g].ssd[r] == mine.ssd[r] && tb[g].ssd[r] < DistanceMetric::UPPER_BOUND &&
                           ids[r] >= 0 && tb[g].o_id[r] < ids[r]);
      if (!better) continue;
      mine.ssd[r] = tb[g].ssd[r];
      ids[r] = tb[g].o_id[r];
      memcpy(&mine.T[12 * r], &tb[g].T[12 * r], 12 * sizeof(float));
      memcpy(&mine.pairs[(size_t)2 * N * r], &tb[g].pairs[(size_t)2 * N * r], sizeof(int16_t) * 2 * (n_min + r));
    }
  TableFromHost(mine, rank);
}

// FirstPass over og[start, stop) (search_or_dyn3d.cc:283-325, 1643-1946).  vector<OrGenEntry> is
// a contiguous array of 12-float Matrix3x4 (or_gen.h:73-78), so the range is passed as is.
void SearchOrCuda::Search(OrientationGenerator const& og, long start, long stop)
{
  const bool alt = settings.num_solutions > 1;
  vector<float> scores;          // [stop - start][R], filled slice by slice by the GPUs
  vector<Real> lowest_before;    // aaSumSqdDists[0] as KeepTrackOfAltSolutions first sees it
  if (alt && alt_method_requires_alt_solution_table) {   // :140-149
    scores.resize((size_t)(stop - start) * R);
    lowest_before.assign(aaSumSqdDists[0], aaSumSqdDists[0] + R);
    reduced_cache.clear();
  }
  SearchRank(0, stop > start ? &og[start].transform[0][0] : NULL, stop - start, start,
             scores.empty() ? NULL : &scores[0], 0, NULL, oid);
  int n;
  for (n = n_min; n <= n_max && aaSumSqdDists[0][n - n_min] == DistanceMetric::UPPER_BOUND; ++n) {}
  found_solutions = n <= n_max;
  if (n > n_max && !fatal_without_solutions) return;   // batch runs: the caller skips this pair
  if (n > n_max)  // search_or_dyn3d.cc:1929-1941
    ERR("Error: No solutions were found satisfying your criteria.\n"
        "       The restrictions you placed on accepting solutions were\n"
        "       too strict.  Try using a lower minimum-number-of-matches,\n"
        "       a larger maximum distance cutoff, for example.");
  if (alt && settings.alt_method == Settings::ALT_METHOD_PAIRS) CalcAlternatesByPairs(og, start, stop, scores, lowest_before);
  else if (alt) CalcAlternates(og, start, stop, scores, lowest_before);
}

// -alt-r (CalcAltSolutionsUsing_ALT_METHOD_PAIRS, search_or_dyn3d.cc:367-416): one more search
// per rank, in which an alignment is recorded only if it shares at most alt_max_pairs_common
// pairs with the solutions of the same size of all lower ranks (MultiPass, :1975-2250 -> one
// mr_search_dissimilar per rank).  With an RMSD tolerance the searches run over the orientations
// that produced, at some size, an alignment within the tolerance of the best one
// (GetOrientationsOfLowRmsdFromAltTable, :960-1035, over the records KeepTrackOfAltSolutions
// keeps: rebuilt from the scores as in CalcAlternates); otherwise over the whole range.
void SearchOrCuda::CalcAlternatesByPairs(OrientationGenerator const& og, long start, long stop,
                                         vector<float> const& scores, vector<Real> const& lowest_before)
{
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  set<long> use_these_orientations;
  if (settings.alt_rmsd_tolerance != Settings::ALTERNATES_NOT_LIMITED_BY_RMSD) {
    vector<vector<float> > cols;
    for (int n = n_min; n <= n_max; ++n) {
      const int row = n - n_min;
      if (row % kScoreBlock == 0) {
        cols.resize(min(kScoreBlock, R - row));
        GatherScoreColumns(scores, stop - start, R, row, cols);
      }
      vector<float> const& column = cols[row % kScoreBlock];
      if (aaSumSqdDists[0][row] == DistanceMetric::UPPER_BOUND) continue;
      const Real best_rmsd = sqrt(aaSumSqdDists[0][row] / (n * nb));
      const Real tolerable_rmsd = settings.alt_rmsd_tolerance * best_rmsd;
      Real lowest_sum_sqd_dist_so_far = lowest_before[row];
      for (long o = start; o < stop; ++o) {
        const Real sum_sqd_dist = column[o - start];
        if (!(sum_sqd_dist < DistanceMetric::UPPER_BOUND)) continue;
        if (sum_sqd_dist < lowest_sum_sqd_dist_so_far) lowest_sum_sqd_dist_so_far = sum_sqd_dist;
        if (!(sum_sqd_dist <= lowest_sum_sqd_dist_so_far * SQR(settings.alt_rmsd_tolerance))) continue;  // :493-497
        const Real rmsd = sqrt(sum_sqd_dist / (n * nb));
        if (tolerable_rmsd >= rmsd) use_these_orientations.insert(o);                                    // :1019-1031
      }
    }
    cout << "  ...After the RMSD-filter, the number of orientations\n"
            "     we have to consider is: " << use_these_orientations.size() << endl;
  } else
    for (long o = start; o < stop; ++o) use_these_orientations.insert(o);
  // the searched list in ascending id order: positions in it break ties like the ids would
  vector<float> list((size_t)12 * use_these_orientations.size());
  size_t at = 0;
  for (set<long>::const_iterator p = use_these_orientations.begin(); p != use_these_orientations.end(); ++p, ++at)
    memcpy(&list[12 * at], &og[*p].transform[0][0], 12 * sizeof(float));
  vector<vector<float> > lower_ssd(settings.num_solutions);
  vector<vector<int16_t> > lower_pairs(settings.num_solutions);
  vector<mr_best> lower(settings.num_solutions);
  for (int r = 1; r < settings.num_solutions; ++r) {
    cout << "----- Searching for the " << r + 1 << "th-best alignments -----" << endl;
    // rank r - 1 is final now: flatten it next to the lower ones
    const int q = r - 1;
    lower_ssd[q].assign(aaSumSqdDists[q], aaSumSqdDists[q] + R);
    lower_pairs[q].assign((size_t)2 * N * R, 0);
    for (int row = 0; row < R; ++row) {
      if (!(lower_ssd[q][row] < DistanceMetric::UPPER_BOUND)) continue;
      for (int k = 0; k < n_min + row; ++k) {
        lower_pairs[q][(size_t)2 * N * row + 2 * k] = aaAlignments[q][row][k][0];
        lower_pairs[q][(size_t)2 * N * row + 2 * k + 1] = aaAlignments[q][row][k][1];
      }
    }
    lower[q].ssd = &lower_ssd[q][0]; lower[q].T = NULL; lower[q].pairs = &lower_pairs[q][0]; lower[q].o_id = NULL;
    vector<int64_t> ids(R, -1);
    SearchRank(r, list.empty() ? NULL : &list[0], (long)use_these_orientations.size(), 0, NULL, r, &lower[0], ids);
    int n;
    for (n = n_min; n <= n_max && aaSumSqdDists[r][n - n_min] == DistanceMetric::UPPER_BOUND; ++n) {}
    if (n > n_max) cout << "No " << r + 1 << "th-best alignments were found." << endl;
  }
}

// The optimal-superposition matrices of every size at orientation o_id (what FirstPass hands
// to KeepTrackOfAltSolutions as orientation_reduced, search_or_dyn3d.cc:1897-1904),
// evaluated on the GPU the first time a candidate of that orientation is looked at.
vector<float> const& SearchOrCuda::ReducedOrientations(OrientationGenerator const& og, long o_id)
{
  map<long, vector<float> >::iterator it = reduced_cache.find(o_id);
  if (it != reduced_cache.end()) return it->second;
  vector<float>& t = reduced_cache[o_id];
  t.resize((size_t)12 * N);
  int32_t n_last = 0, nw = 0;
  Check(ctx[0], mr_eval_one(ctx[0], &og[o_id].transform[0][0], 0, &n_last, &nw, NULL, NULL, NULL, &t[0], NULL));
  return t;
}

// ReadAltSolutionFiles + CalcAltSolutionsFromAltSolutionTable (search_or_dyn3d.cc:626-920,
// 1253-1447) without the temp files.  The file of size n holds one record per orientation
// whose score was within alt_rmsd_tolerance of the best score SO FAR (KeepTrackOfAltSolutions,
// :477-542), in search order; that sequence is rebuilt here from the scores of the whole
// search and a running minimum, then sorted and walked exactly as ReadAltSolutionFile does
// (same element type, same std::sort, same comparisons in Real).
void SearchOrCuda::CalcAlternates(OrientationGenerator const& og, long start, long stop,
                                  vector<float> const& scores, vector<Real> const& lowest_before)
{
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  const bool m3d = settings.alt_method == Settings::ALT_METHOD_3D;
  const bool unlimited = settings.alt_rmsd_tolerance == Settings::ALTERNATES_NOT_LIMITED_BY_RMSD;
  cout << "Choosing the best alternate alignments from the scores of the search." << endl;
  multimap<long, OrientationOwner> orientation_owners;
  vector<long> alt_ids(settings.num_solutions);
  vector<float> alt_mats((size_t)12 * settings.num_solutions);
  vector<vector<float> > cols;
  for (int n = n_min; n <= n_max; ++n) {
    const int row = n - n_min;
    if (row % kScoreBlock == 0) {
      cols.resize(min(kScoreBlock, R - row));
      GatherScoreColumns(scores, stop - start, R, row, cols);
    }
    vector<float> const& column = cols[row % kScoreBlock];
    if (aaSumSqdDists[0][row] == DistanceMetric::UPPER_BOUND) continue;  // :641
    vector<OrientationUsageInstance> vInstances;
    OrientationUsageInstance best_orientation;
    bool best_orientation_found = false;
    Real lowest_sum_sqd_dist_so_far = lowest_before[row];
    for (long o = start; o < stop; ++o) {
      const Real sum_sqd_dist = column[o - start];
      if (!(sum_sqd_dist < DistanceMetric::UPPER_BOUND)) continue;  // never compared at this size
      if (sum_sqd_dist < lowest_sum_sqd_dist_so_far) lowest_sum_sqd_dist_so_far = sum_sqd_dist;  // :1825-1835
      if (!((sum_sqd_dist <= lowest_sum_sqd_dist_so_far * SQR(settings.alt_rmsd_tolerance)) || unlimited))
        continue;                                                    // :493-501
      OrientationUsageInstance u;
      u.orig_orientation_id = o;
      u.rmsd = sqrt(sum_sqd_dist / (n * nb));                        // :527-529
      // :722-746: the record of the stored best solution goes first; it is the record of the
      // winning orientation (the first one, in search order, that reached the best score)
      if (!best_orientation_found && o == oid[row]) {
        best_orientation = u;
        best_orientation_found = true;
      } else
        vInstances.push_back(u);
    }
    if (!best_orientation_found)
      ERR_INTERNAL("alternates: the best solution with " << n << " matches was not found in this search range.");
    sort(vInstances.begin(), vInstances.end());                      // :755
    vInstances.insert(vInstances.begin(), best_orientation);
    memcpy(&vInstances[0].orientation_reduced[0][0], &aaTransforms[0][row][0][0], 12 * sizeof(Real));
    alt_ids[0] = vInstances[0].orig_orientation_id;
    memcpy(&alt_mats[0], &aaTransforms[0][row][0][0], 12 * sizeof(Real));
    const Real best_rmsd = vInstances[0].rmsd;
    int alt_count = 1;
    for (vector<OrientationUsageInstance>::iterator p = vInstances.begin() + 1;
         p != vInstances.end() && alt_count < settings.num_solutions &&
         ((p->rmsd <= best_rmsd * settings.alt_rmsd_tolerance) || unlimited);
         ++p) {
      if (m3d)
        memcpy(&p->orientation_reduced[0][0], &ReducedOrientations(og, p->orig_orientation_id)[(size_t)12 * (n - 1)],
               12 * sizeof(Real));
      int r;
      for (r = 0; r < alt_count &&
                  SufficientlyDifferentOrientationFromRthAlternate(
                      *p, &alt_ids[0], reinterpret_cast<Matrix3x4*>(&alt_mats[0]), r, og);
           ++r) {}
      if (r != alt_count) continue;                                  // :858
      OrientationOwner entry;
      entry.n = n;
      entry.r = alt_count;
      orientation_owners.insert(make_pair(p->orig_orientation_id, entry));
      alt_ids[alt_count] = p->orig_orientation_id;
      if (m3d) memcpy(&alt_mats[(size_t)12 * alt_count], &p->orientation_reduced[0][0], 12 * sizeof(Real));
      ++alt_count;
    }
  }
  // :1253-1447: every owning orientation is evaluated once more (on the GPU) and the sizes it
  // owns are stored at their rank
  cout << "Recalculating the candidates for alternate alignments on the GPU." << endl;
  vector<float> cost(N), ssd_n(N), T_n((size_t)12 * N);
  vector<int16_t> pairs_n((size_t)2 * N * N);
  vector<uint8_t> valid_n(N);
  int orientation_counter = 0, alignment_counter = 0;
  multimap<long, OrientationOwner>::iterator q = orientation_owners.begin();
  while (q != orientation_owners.end()) {
    const long o_id = q->first;
    int32_t n_last = 0, nw = 0;
    Check(ctx[0], mr_eval_one(ctx[0], &og[o_id].transform[0][0], 0, &n_last, &nw, &cost[0], &valid_n[0], &ssd_n[0],
                              &T_n[0], &pairs_n[0]));
    for (; q != orientation_owners.end() && q->first == o_id; ++q) {
      const int n = q->second.n, r = q->second.r, row = n - n_min;
      if (!valid_n[n - 1]) ERR_INTERNAL("alternates: orientation " << o_id << " has no alignment with " << n << " matches.");
      // :1404-1423
      aaSumSqdDists[r][row] = settings.refine_while_searching ? ssd_n[n - 1] : cost[n - 1];
      if (settings.refine_while_searching)
        memcpy(&aaTransforms[r][row][0][0], &T_n[(size_t)12 * (n - 1)], 12 * sizeof(Real));
      else
        CopyMatrix3x4(og[o_id].transform, aaTransforms[r][row]);
      aaAlignments[r][row].resize(n);
      for (int k = 0; k < n; ++k) {
        aaAlignments[r][row][k][0] = pairs_n[(size_t)2 * N * (n - 1) + 2 * k];
        aaAlignments[r][row][k][1] = pairs_n[(size_t)2 * N * (n - 1) + 2 * k + 1];
      }
      ++alignment_counter;
    }
    ++orientation_counter;
  }
  cout << "During re-calculation to find the alternate alignments,\n"
       << orientation_counter << " orientations were tried, and\n"
       << alignment_counter << " alternate alignments were found\n" << flush;
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
  if (settings.num_solutions > 1 && settings.refine_method == Settings::REFINE_ALL) {
    int num_iterations = 0, num_new_solutions = 0;
    for (int r = 1; r < settings.num_solutions; ++r) RefineAlternate(r, num_iterations, num_new_solutions);
    cout << "(Refining the alternates, " << num_iterations << " new orientations were tested on the GPU,\n"
            " and " << num_new_solutions << " instances of better alignments were found.)" << endl;
  }
  // :2648-2654: the refined solutions may have moved next to one another (the reference's
  // own table bookkeeping, a handful of O(numRes2) comparisons per size)
  if (settings.num_solutions > 1) CheckAlternatesForConsistency();
}

// Refine() for a rank r > 0 (-r-all, search_or_dyn3d.cc:2344-2640): the size n at the median of
// the unrefined ones is superimposed on its own alignment (CalcMinRMSD,
// lib/pair_alignment/pair_alignment.cc:715-820: pairs in order, the nb atoms of a pair
// consecutively), the DP runs at that orientation, and -- unlike rank 0 -- only the solution of
// the SAME size may be replaced (:2543-2560; the m-loop stops at m = n, :2415-2421).  A size is
// done when an evaluation leaves its alignment unchanged.  Both steps run on the GPU
// (mr_superimpose_batch, mr_eval_one); valid[n-1] says that none of the loop's early exits
// (max-dist :2424, max-rmsd :2500) fired at or before n.
void SearchOrCuda::RefineAlternate(int r, int& num_iterations_total, int& num_new_solutions)
{
  const int nb = Biopolymer::Residue::NumBackboneAtoms();
  vector<char> flags(R);
  bool* aNeed2Refine = reinterpret_cast<bool*>(&flags[0]);
  for (int row = 0; row < R; ++row) aNeed2Refine[row] = aaSumSqdDists[r][row] < DistanceMetric::UPPER_BOUND;
  cout << "Currently refining the " << r + 1 << "th-best alignments.\n" << flush;
  vector<float> fixed((size_t)3 * nb * N), moveable((size_t)3 * nb * N), cost(N), ssd_n(N), T_n((size_t)12 * N);
  vector<int16_t> pairs_n((size_t)2 * N * N);
  vector<uint8_t> valid_n(N);
  PairwiseAlignment oldNthAlignment(max_num_matches);
  for (int num_iterations = 1;
       num_iterations <= settings.refine_max_iters || settings.refine_max_iters == Settings::ITERATE_UNTIL_CONVERGENCE;
       ++num_iterations) {
    const int n = FindMedian(aNeed2Refine, n_min, n_max);
    if (n == n_min - 1) break;
    ++num_iterations_total;
    const int row = n - n_min;
    PairwiseAlignment& al = aaAlignments[r][row];
    for (int k = 0; k < n; ++k)
      for (int q = 0; q < nb; ++q) {
        memcpy(&fixed[(size_t)3 * (k * nb + q)], &coordsA[(size_t)3 * (al[k][0] * nb + q)], 3 * sizeof(float));
        memcpy(&moveable[(size_t)3 * (k * nb + q)], &coordsB[(size_t)3 * (al[k][1] * nb + q)], 3 * sizeof(float));
      }
    float T12[12], rmsd = 0.0f;
    Check(ctx[0], mr_superimpose_batch(ctx[0], 1, n * nb, &fixed[0], &moveable[0], T12, &rmsd));
    int32_t n_last = 0, nw = 0;
    Check(ctx[0], mr_eval_one(ctx[0], T12, 0, &n_last, &nw, &cost[0], &valid_n[0], &ssd_n[0], &T_n[0], &pairs_n[0]));
    oldNthAlignment.resize(n);
    oldNthAlignment = al;
    if (valid_n[n - 1]) {
      const Real sum_sqd_dist = settings.refine_while_searching ? ssd_n[n - 1] : cost[n - 1];
      if (sum_sqd_dist < aaSumSqdDists[r][row]) {
        aaSumSqdDists[r][row] = sum_sqd_dist;
        memcpy(&aaTransforms[r][row][0][0], settings.refine_while_searching ? &T_n[(size_t)12 * (n - 1)] : T12,
               12 * sizeof(Real));
        for (int k = 0; k < n; ++k) {
          al[k][0] = pairs_n[(size_t)2 * N * (n - 1) + 2 * k];
          al[k][1] = pairs_n[(size_t)2 * N * (n - 1) + 2 * k + 1];
        }
        ++num_new_solutions;
      }
    }
    aNeed2Refine[row] = !(oldNthAlignment == al);
  }
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
  // one pair of files per rank, named align_ / align<r+1>_ as at :2682-2688, 2968-2990
  for (int r = 0; r < settings.num_solutions; ++r) {
    WritePlotFileCuda(ca, r);
    stringstream base, rank;
    base << SearchOrDyn3d::BASE_FILE_NAME;
    if (settings.num_solutions != 1) base << r + 1;
    string info_name = base.str() + "_chimera.info";
    ofstream info_file(info_name.c_str(), ios::out);
    if (!info_file) ERR_INTERNAL("Cannot open file \"" << info_name << "\" for writing.  Exiting...");
    info_file << "fixed pdb:" << settings.vPdbFileNames[0] << "\n";
    info_file << "moveable pdb:" << settings.vPdbFileNames[1] << "\n";
    info_file << "rmsd:" << base.str() << "_chimera.plot\n";
    info_file << "base:" << base.str() << (settings.num_solutions != 1 ? "_" : "") << "\n";
    info_file << "number of atoms used:" << nb << "\n";
    for (int q = 0; q < nb; ++q) info_file << Biopolymer::Residue::LookupBackboneAtomSymbol(q) << "\n";
  }
}

void SearchOrCuda::WritePlotFileCuda(int ca_index, int rank) const
{
  vector<float> vs(R), vT(12 * R);
  vector<int16_t> vp((size_t)2 * N * R, 0);
  for (int r = 0; r < R; ++r) {
    vs[r] = aaSumSqdDists[rank][r];
    memcpy(&vT[12 * r], &aaTransforms[rank][r][0][0], 12 * sizeof(float));
    if (!(vs[r] < DistanceMetric::UPPER_BOUND)) continue;
    for (int k = 0; k < n_min + r; ++k) {
      vp[(size_t)2 * N * r + 2 * k] = aaAlignments[rank][r][k][0];
      vp[(size_t)2 * N * r + 2 * k + 1] = aaAlignments[rank][r][k][1];
    }
  }
  mr_best t;
  t.ssd = &vs[0]; t.T = &vT[0]; t.pairs = &vp[0]; t.o_id = NULL;
  vector<double> rows((size_t)9 * R);
  Check(ctx[0], mr_plot_stats(ctx[0], &t, ca_index, &rows[0]));
  stringstream plot_base;
  plot_base << SearchOrDyn3d::BASE_FILE_NAME;
  if (settings.num_solutions != 1) plot_base << rank + 1;
  string plot_name = plot_base.str() + "_chimera.plot";
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
