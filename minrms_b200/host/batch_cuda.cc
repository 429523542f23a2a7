// batch_cuda.cc -- all-vs-all driver behind the reference's own command line (SURVEY.md 8f rank 2,
// BASELINE config 5).  The reference has no batch mode; every unordered pair of a set of
// structures is an independent `minrms [flags] A B` run (bin/minrms/main.cc:60-176: parse,
// OrientationGenerator, Search(og, 0, og.size()), Refine(), WriteMSF(), WriteChimeraInfo()), so
// the pairs shard over GPUs with no exchange.  The reference's parser keeps process-wide state
// that does not survive a second parse (Biopolymer::Residue::StoreBackboneAtomName's static
// atom table, lib/biopolymer/biopolymer.cc:163-200; DistanceMetric::max_allowed_dist_metric), so
// every pair runs main.cc's sequence with SearchOrCuda in a PROCESS of its own, forked by a
// parent that never touches CUDA; G x K slots (K per GPU) are kept busy by list scheduling, the
// pairs taken largest first (cost = fragment pairs x Dyn3d cells, as in minrms_b200/batch.py).
// K > 1 overlaps one pair's start-up (CUDA context, PDB parsing) with another pair's search.
// The files of pair (A, B) are what `minrms [flags] A B` writes, in <out>/<A>__<B>/.
//
//   minrms_batch_cuda [minrms flags ...] -batch LIST [-batch-out DIR] [-batch-gpus G]
//                     [-batch-workers-per-gpu K]
//   LIST: text file, one structure argument per line (a PDB path, optionally with the
//         reference's ",selection" suffix).  Flags are passed to every pair unchanged.
//
// Nothing here computes: orientation generation on the host is the reference's own, everything
// from the search on runs in libminrms_b200.so.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>
using namespace std;

#include <global_utils.h>
#include <parse_utils.h>
#include <or_gen.h>
#include "minrms_parser.h"
#include "search_or_cuda.h"

extern const char g_version_string[] = "1.0.1-b200";
extern const char g_date_string[] = "<cuda batch>";

using namespace minrms;

namespace {

struct Pair { int i, j; double cost; };

string BaseName(string const& arg)
{
  string s = arg.substr(0, arg.find(','));            // drop the ",selection" suffix
  size_t slash = s.find_last_of('/');
  if (slash != string::npos) s = s.substr(slash + 1);
  size_t dot = s.find_last_of('.');
  if (dot != string::npos && dot > 0) s = s.substr(0, dot);
  return s;
}

string Absolute(string const& arg, string const& cwd)
{
  return (!arg.empty() && arg[0] == '/') ? arg : cwd + "/" + arg;
}

// residues of a structure for the schedule: C-alpha ATOM records of the file (selections and
// atoms_used.txt may change the count a little; the schedule only needs relative sizes)
long CountResidues(string const& arg, string const& cwd)
{
  ifstream f(Absolute(arg.substr(0, arg.find(',')), cwd).c_str());
  string line;
  long n = 0;
  while (getline(f, line))
    if (line.size() > 15 && line.compare(0, 4, "ATOM") == 0 && line.compare(12, 4, " CA ") == 0) ++n;
  return n;
}

// cell updates of one pair: fragment pairs x Dyn3d cells per orientation (closed form of
// SURVEY.md 8 a7 through the library)
double PairCost(long m, long n, int frag)
{
  if (m < frag || n < frag) return 0.0;
  const long N = min(m, n);
  return (double)(m - frag + 1) * (double)(n - frag + 1) * (double)mr_dp_cells_per_orientation((int)m, (int)n, 1, (int)N);
}

double Now()
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

// one pair = main.cc:105-173 with SearchOrCuda; files go to out_dir
void RunPair(vector<string> const& flags, string const& a, string const& b, string const& out_dir,
             string const& cwd)
{
  vector<string> vArgs(flags);
  vArgs.push_back(a);
  vArgs.push_back(b);
  if (chdir(cwd.c_str()) != 0) ERR("Error: cannot change to \"" << cwd << "\"");  // atoms_used.txt etc. are read from here
  MinrmsParser settings(vArgs);
  if (vArgs.size() != 1) ERR("Error: Unrecognized command line option: \"" << vArgs[1] << "\"\n");
  OrientationGenerator og(settings);
  SearchOrCuda search(settings);
  search.SetFatalWithoutSolutions(false);
  search.Search(og, 0, og.size());
  mkdir(out_dir.c_str(), 0777);
  if (chdir(out_dir.c_str()) != 0) ERR("Error: cannot write to \"" << out_dir << "\"");
  if (!search.FoundSolutions()) {   // `minrms A B` would have ended with "No solutions were found"
    ofstream(string("NO_SOLUTIONS").c_str()) << "No solutions were found satisfying your criteria.\n";
    return;
  }
  if (settings.refine_method != SearchOr::Settings::NO_REFINE) search.Refine();
  search.WriteMSF();
  search.WriteChimeraInfo();
}

}  // namespace

int main(int argc, char** argv)
{
  srand48(42);  // main.cc:183-195
  vector<string> vOrig, vArgs;
  ParseUtils::ConvertArgvToVectorOfStrings(argc, argv, vOrig);
  char back_slash[2] = {'\\', '\0'};
  const char* dflt = getenv("MINRMS_SETTINGS_FILE");
  ParseUtils::HandleArgFiles(vOrig, vArgs, "-cfg", "#", back_slash, dflt ? string(dflt) : string());
  string list_name, out = "batch_out";
  int gpus = 1, per_gpu = 1, frag = 4;
  vector<string> flags;
  for (size_t k = 0; k < vArgs.size(); ++k) {
    const bool has_next = k + 1 < vArgs.size();
    if (vArgs[k] == "-batch" && has_next) list_name = vArgs[++k];
    else if (vArgs[k] == "-batch-out" && has_next) out = vArgs[++k];
    else if (vArgs[k] == "-batch-gpus" && has_next) gpus = max(1, atoi(vArgs[++k].c_str()));
    else if (vArgs[k] == "-batch-workers-per-gpu" && has_next) per_gpu = max(1, atoi(vArgs[++k].c_str()));
    else {
      if (vArgs[k] == "-fm" && has_next) frag = max(1, atoi(vArgs[k + 1].c_str()));
      flags.push_back(vArgs[k]);
    }
  }
  if (list_name.empty())
    ERR("usage: minrms_batch_cuda [minrms flags] -batch LIST [-batch-out DIR] [-batch-gpus G] [-batch-workers-per-gpu K]");
  char cwd_buf[4096];
  if (!getcwd(cwd_buf, sizeof(cwd_buf))) ERR("Error: getcwd failed");
  const string cwd(cwd_buf);
  vector<string> structures;
  {
    ifstream f(list_name.c_str());
    if (!f) ERR("Error: cannot read \"" << list_name << "\"");
    string line;
    while (getline(f, line)) {
      while (!line.empty() && isspace((unsigned char)line[line.size() - 1])) line.erase(line.size() - 1);
      if (!line.empty() && line[0] != '#') structures.push_back(line);
    }
  }
  if (structures.size() < 2) ERR("Error: \"" << list_name << "\" names fewer than two structures.");
  const string out_abs = Absolute(out, cwd);
  mkdir(out_abs.c_str(), 0777);

  // the schedule: all unordered pairs, largest first, each to the least loaded worker
  vector<long> len(structures.size());
  for (size_t s = 0; s < structures.size(); ++s) len[s] = CountResidues(structures[s], cwd);
  vector<Pair> pairs;
  for (int i = 0; i < (int)structures.size(); ++i)
    for (int j = i + 1; j < (int)structures.size(); ++j) {
      Pair p = {i, j, PairCost(len[i], len[j], frag)};
      pairs.push_back(p);
    }
  stable_sort(pairs.begin(), pairs.end(), [](Pair const& x, Pair const& y) { return x.cost > y.cost; });
  const int W = gpus * per_gpu;
  cout << "minrms batch v" << g_version_string << ": " << structures.size() << " structures, " << pairs.size()
       << " pairs, " << W << " slot(s) on " << gpus << " GPU(s)\n" << flush;

  // List scheduling, largest pair first: whenever one of the W slots (slot w runs on GPU
  // w % gpus) is free, the next pair of the list starts there in a process of its own.
  const double t0 = Now();
  vector<pid_t> slot_pid(W, (pid_t)0);
  vector<size_t> slot_pair(W, 0);
  size_t next = 0, running = 0;
  int failed = 0;
  while (next < pairs.size() || running > 0) {
    for (int w = 0; w < W && next < pairs.size(); ++w) {
      if (slot_pid[w] != 0) continue;
      Pair const& p = pairs[next];
      const string dir = out_abs + "/" + BaseName(structures[p.i]) + "__" + BaseName(structures[p.j]);
      fflush(stdout);
      const pid_t pid = fork();
      if (pid < 0) ERR("Error: fork failed");
      if (pid == 0) {
        // the child's GPU: set before the first CUDA call of this process
        char dev[16];
        snprintf(dev, sizeof(dev), "%d", w % gpus);
        if (gpus > 1 || !getenv("CUDA_VISIBLE_DEVICES")) setenv("CUDA_VISIBLE_DEVICES", dev, 1);
        unsetenv("MINRMS_CUDA_GPUS");
        mkdir(dir.c_str(), 0777);
        FILE* log = freopen((dir + "/minrms.log").c_str(), "w", stdout);   // the reference's progress output
        (void)log;
        dup2(fileno(stdout), 2);   // and its warnings
        RunPair(flags, structures[p.i], structures[p.j], dir, cwd);
        fflush(stdout);
        _exit(0);
      }
      slot_pid[w] = pid;
      slot_pair[w] = next++;
      ++running;
    }
    int status = 0;
    const pid_t done = waitpid(-1, &status, 0);
    if (done <= 0) break;
    for (int w = 0; w < W; ++w) {
      if (slot_pid[w] != done) continue;
      Pair const& p = pairs[slot_pair[w]];
      if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        ++failed;
        cerr << "pair " << BaseName(structures[p.i]) << " " << BaseName(structures[p.j]) << " FAILED ("
             << (WIFSIGNALED(status) ? "signal " : "exit status ")
             << (WIFSIGNALED(status) ? WTERMSIG(status) : WEXITSTATUS(status)) << "), see its minrms.log\n";
      }
      slot_pid[w] = 0;
      --running;
    }
  }
  const double dt = Now() - t0;
  cout << pairs.size() << " pairs in " << dt << " s (" << pairs.size() / dt << " pairs/s)";
  if (failed) cout << "; " << failed << " pair(s) FAILED";
  cout << endl;
  return failed ? 1 : 0;
}
