// batch_cuda.cc -- all-vs-all driver behind the reference's own command line (SURVEY.md 8f rank 2,
// BASELINE config 5).  The reference has no batch mode; every unordered pair of a set of
// structures is an independent `minrms [flags] A B` run (bin/minrms/main.cc:60-176: parse,
// OrientationGenerator, Search(og, 0, og.size()), Refine(), WriteMSF(), WriteChimeraInfo()), so
// the pairs shard over GPUs with no exchange: G x K worker processes (K per GPU, forked before any
// CUDA call; K > 1 overlaps one pair's parsing, host-side fragment matching and Refine with
// another pair's search), each taking a cost-balanced share of the pair list (largest first, to
// the least loaded worker: the rule of minrms_b200/batch.py; cost = fragment pairs x Dyn3d cells)
// and running main.cc's sequence with SearchOrCuda pair after pair -- the CUDA context and the
// library's parked work space are reused from one pair to the next.
// The reference's parser was written for one run per process: Biopolymer::Residue keeps the
// names of the atoms used in a static table (lib/biopolymer/biopolymer.cc:166-200) to which a
// second parse appends, shifting the atom ids; the table is emptied before every parse (see
// BackboneTable below) once the previous pair's objects are gone.  ERR() ends a process
// (ERR.h:20-21), so a pair the reference refuses takes its worker down: the parent notices,
// reports the pair as failed and starts a new worker on the rest of that share.
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

// Access to the private static Biopolymer::Residue::backboneAtomLookup without touching the
// reference's headers: an explicit template instantiation may name a private member
// ([temp.spec]/6), and the friend function it defines hands the address out.
struct BackboneTable {
  typedef map<string, int>* type;
  friend type get(BackboneTable);
};
template <typename Tag, typename Tag::type Member>
struct Expose {
  friend typename Tag::type get(Tag) { return Member; }
};
template struct Expose<BackboneTable, &Biopolymer::Residue::backboneAtomLookup>;

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
  get(BackboneTable())->clear();   // the parser registers the atoms used again, from id 0
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
  vector<double> load(W, 0.0);
  vector<vector<Pair> > share(W);
  for (size_t k = 0; k < pairs.size(); ++k) {
    int w = 0;
    for (int c = 1; c < W; ++c) if (load[c] < load[w]) w = c;
    share[w].push_back(pairs[k]);
    load[w] += pairs[k].cost;
  }
  cout << "minrms batch v" << g_version_string << ": " << structures.size() << " structures, " << pairs.size()
       << " pairs, " << W << " worker(s) on " << gpus << " GPU(s)\n" << flush;

  const double t0 = Now();
  vector<pid_t> worker_pid(W, (pid_t)0);
  vector<size_t> worker_from(W, 0);   // first pair of its share the current process of worker w was given
  int failed = 0;
  size_t running = 0;
  // worker w = GPU w % gpus; runs share[w][from...] in order, leaving <dir>/minrms.done behind each pair
  auto pair_dir = [&](Pair const& p) { return out_abs + "/" + BaseName(structures[p.i]) + "__" + BaseName(structures[p.j]); };
  auto spawn = [&](int w, size_t from) {
    fflush(stdout);
    const pid_t pid = fork();
    if (pid < 0) ERR("Error: fork failed");
    if (pid == 0) {
      char dev[16];
      snprintf(dev, sizeof(dev), "%d", w % gpus);   // set before the first CUDA call of this process
      if (gpus > 1 || !getenv("CUDA_VISIBLE_DEVICES")) setenv("CUDA_VISIBLE_DEVICES", dev, 1);
      unsetenv("MINRMS_CUDA_GPUS");
      for (size_t k = from; k < share[w].size(); ++k) {
        Pair const& p = share[w][k];
        const string dir = pair_dir(p);
        mkdir(dir.c_str(), 0777);
        FILE* log = freopen((dir + "/minrms.log").c_str(), "w", stdout);   // the reference's progress output
        (void)log;
        dup2(fileno(stdout), 2);                                            // and its warnings
        RunPair(flags, structures[p.i], structures[p.j], dir, cwd);
        fflush(stdout);
        ofstream((dir + "/minrms.done").c_str()) << "done\n";
      }
      _exit(0);
    }
    worker_pid[w] = pid;
    worker_from[w] = from;
    ++running;
  };
  for (int w = 0; w < W; ++w)
    if (!share[w].empty()) spawn(w, 0);
  while (running > 0) {
    int status = 0;
    const pid_t done = waitpid(-1, &status, 0);
    if (done <= 0) break;
    for (int w = 0; w < W; ++w) {
      if (worker_pid[w] != done) continue;
      worker_pid[w] = 0;
      --running;
      if (WIFEXITED(status) && WEXITSTATUS(status) == 0) continue;
      // the first pair without its marker is the one that ended the process
      size_t k = worker_from[w];
      while (k < share[w].size() && ifstream((pair_dir(share[w][k]) + "/minrms.done").c_str())) ++k;
      if (k < share[w].size()) {
        Pair const& p = share[w][k];
        ++failed;
        cerr << "pair " << BaseName(structures[p.i]) << " " << BaseName(structures[p.j]) << " FAILED ("
             << (WIFSIGNALED(status) ? "signal " : "exit status ")
             << (WIFSIGNALED(status) ? WTERMSIG(status) : WEXITSTATUS(status)) << "), see its minrms.log\n";
        if (k + 1 < share[w].size()) spawn(w, k + 1);
      }
    }
  }
  const double dt = Now() - t0;
  cout << pairs.size() << " pairs in " << dt << " s (" << pairs.size() / dt << " pairs/s)";
  if (failed) cout << "; " << failed << " pair(s) FAILED";
  cout << endl;
  return failed ? 1 : 0;
}
