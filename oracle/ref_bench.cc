// ref_bench.cc -- TEST/BENCH INFRASTRUCTURE.  Built only where /root/reference exists
// (oracle/build_ref.sh), linked against the UNMODIFIED reference objects.
//
// Usage: ref_bench PROCS PER_PROC REPEAT <minrms arguments...>
//
// Times the reference's own SearchOrDyn3d::Search(og, start, stop)
// (bin/minrms/search_or.h:87-104 -- the range API the author left "to help anyone who
// tries to write a parallel-ized version") on PROCS forked processes, each over its own
// disjoint range of PER_PROC orientations of the reference's shuffled list.  Prints one
// line per repeat:   ref_bench procs=P orientations=K seconds=S
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <iostream>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>
using namespace std;

#include <global_utils.h>
#include <parse_utils.h>
#include <or_gen.h>
#include "search_or.h"
#include "search_or_dyn3d.h"
#include "minrms_parser.h"

using namespace minrms;
extern const char g_version_string[] = "ref_bench";
extern const char g_date_string[] = "<oracle>";

static double now() { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + 1e-6 * t.tv_usec; }

int main(int argc, char **argv)
{
  if (argc < 5) { fprintf(stderr, "usage: ref_bench PROCS PER_PROC REPEAT <minrms args>\n"); return 2; }
  int procs = atoi(argv[1]);
  long per = atol(argv[2]);
  int repeat = atoi(argv[3]);
  vector<string> vOrig;
  vOrig.push_back("minrms");
  for (int i = 4; i < argc; i++) vOrig.push_back(argv[i]);
  char back_slash[2] = {'\\', '\0'};
  vector<string> vArgs;
  int saved = dup(1);
  freopen("/dev/null", "w", stdout);  // the reference prints progress text
  ParseUtils::HandleArgFiles(vOrig, vArgs, "-cfg", "#", back_slash, string());
  MinrmsParser settings(vArgs);
  OrientationGenerator og(settings);
  long total = og.size();
  if (per * procs > total) per = total / procs;
  for (int rep = 0; rep < repeat; rep++) {
    double t0 = now();
    vector<pid_t> kids;
    for (int p = 0; p < procs; p++) {
      pid_t pid = fork();
      if (pid == 0) {
        SearchOrDyn3d s(settings);
        s.Search(og, p * per, (p + 1) * per);
        _exit(0);
      }
      kids.push_back(pid);
    }
    int bad = 0;
    for (size_t k = 0; k < kids.size(); k++) { int st; waitpid(kids[k], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) bad++; }
    double t1 = now();
    dprintf(saved, "ref_bench procs=%d orientations=%ld total=%ld seconds=%.6f failed=%d\n", procs, per * procs, total, t1 - t0, bad);
  }
  return 0;
}
