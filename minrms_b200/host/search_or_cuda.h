// search_or_cuda.h -- the reference-side binding of the B200 search: a SearchOr
// (bin/minrms/search_or.h:57-130) whose Search() and Refine() run on the GPU through the
// C ABI of include/minrms_b200.h.  It derives from SearchOrDyn3d only to reuse the
// reference's own result tables and writers (WriteMSF / WriteChimeraInfo,
// bin/minrms/search_or_dyn3d.cc:2848-2996) unchanged.  Alternate solutions: -alt-o / -alt-o-orig
// (the scores of FirstPass from the GPU, the reference's selection rules on the host), -alt-r
// (one more GPU search per rank with the common-pairs test inside the traceback kernel) and
// -r-all (the alternates refined through the GPU as well).
//
// SearchOrDyn3d keeps aaSumSqdDists / aaTransforms / aaAlignments private
// (bin/minrms/search_or_dyn3d.h:228-310).  A maintainer adopting this class changes that
// one access specifier to `protected`; to build against the UNMODIFIED tree (as
// oracle/build_ref.sh does) the specifier is redefined around the include below.
#ifndef MINRMS_SEARCH_OR_CUDA_H
#define MINRMS_SEARCH_OR_CUDA_H

// Everything search_or_dyn3d.h includes is included first, under the keywords' own meaning
// (their include guards then skip them), so that the redefinition reaches the declaration
// of SearchOrDyn3d and nothing else -- no standard or library header is parsed with it.
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <sys/time.h>
#include <biopolymer.h>
#include <pair_alignment.h>
#include <superimpose.h>
#include <dyn3d.h>
#include "search_or.h"
#include "or_gen.h"

#define private protected
#include "search_or_dyn3d.h"
#undef private

#include "minrms_b200.h"

namespace minrms {

class SearchOrCuda : public SearchOrDyn3d {
 public:
  explicit SearchOrCuda(SearchOrDyn3d::Settings const& s);
  ~SearchOrCuda();

  // SearchOr interface (bin/minrms/search_or.h:87-114)
  void Search(OrientationGenerator const& og, long start, long stop);
  void Refine();
  // SearchOr::WriteChimeraInfo (search_or.h:125-129): align_chimera.info as the reference
  // writes it and align_chimera.plot with its per-size statistics computed on the GPU
  void WriteChimeraInfo() const;
  // The reference ends the process when a search finds nothing (search_or_dyn3d.cc:1929-1941);
  // a batch driver turns that off and asks afterwards.
  void SetFatalWithoutSolutions(bool fatal) { fatal_without_solutions = fatal; }
  bool FoundSolutions() const { return found_solutions; }

 private:
  bool fatal_without_solutions = true;
  bool found_solutions = false;
  SearchOrCuda(SearchOrCuda const&);
  SearchOrCuda& operator=(SearchOrCuda const&);
  void TableToHost(mr_best& t, int rank = 0);   // reference tables of a rank -> flat C-ABI table
  void TableFromHost(mr_best const& t, int rank = 0);
  void SearchRank(int rank, const float* first12, long count, long id0, float* scores, int n_lower,
                  const mr_best* lower, std::vector<int64_t>& ids);
  // -alt-r (ALT_METHOD_PAIRS): search_or_dyn3d.cc:367-416, 960-1035, 1975-2250
  void CalcAlternatesByPairs(OrientationGenerator const& og, long start, long stop,
                             std::vector<float> const& scores, std::vector<Real> const& lowest_before);
  void WritePlotFileCuda(int ca_index, int r) const;
  // -alt-o / -alt-o-orig (ALT_METHOD_3D / ALT_METHOD_3D_ORIG): the reference's bookkeeping
  // of search_or_dyn3d.cc:477-542, 626-952, 1253-1447 from the scores of mr_search_scores
  void CalcAlternates(OrientationGenerator const& og, long start, long stop,
                      std::vector<float> const& scores, std::vector<Real> const& lowest_before);
  std::vector<float> const& ReducedOrientations(OrientationGenerator const& og, long o_id);
  // -r-all: the ranks above the first (search_or_dyn3d.cc:2344-2640 with r > 0)
  void RefineAlternate(int r, int& num_iterations, int& num_new_solutions);
  std::vector<float> coordsA, coordsB;  // the constructor's flattened atoms (CalcMinRMSD overwrites the base's arrays)
  std::map<long, std::vector<float> > reduced_cache;  // o_id -> mr_eval_one's T, [N][12]
  std::vector<mr_ctx*> ctx;       // one per GPU ($MINRMS_CUDA_GPUS, default 1)
  std::vector<float> ssd, T;
  std::vector<int16_t> pairs;
  std::vector<int64_t> oid;
  int N, R;
};

}  // namespace minrms
#endif
