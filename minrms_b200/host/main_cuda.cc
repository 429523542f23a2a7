// main_cuda.cc -- the reference's command line (bin/minrms/main.cc:60-176 flow: parse,
// generate orientations, search, refine, write) with SearchOrCuda in the place of
// SearchOrDyn3d.  Flags, PDB loading, orientation generation and the MSF / Chimera
// writers are the reference's own code, linked unmodified.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
using namespace std;

#include <global_utils.h>
#include <parse_utils.h>
#include <or_gen.h>
#include "minrms_parser.h"
#include "search_or_cuda.h"

extern const char g_version_string[] = "1.0.1-b200";
extern const char g_date_string[] = "<cuda>";

using namespace minrms;

int main(int argc, char** argv)
{
  cout << "minrms v" << g_version_string << " " << g_date_string << "\n" << flush;
  srand48(42);  // main.cc:183-195
  vector<string> vOrig, vArgs;
  ParseUtils::ConvertArgvToVectorOfStrings(argc, argv, vOrig);
  char back_slash[2] = {'\\', '\0'};
  const char* dflt = getenv("MINRMS_SETTINGS_FILE");
  ParseUtils::HandleArgFiles(vOrig, vArgs, "-cfg", "#", back_slash, dflt ? string(dflt) : string());
  MinrmsParser settings(vArgs);
  if (vArgs.size() != 1) ERR("Error: Unrecognized command line option: \"" << vArgs[1] << "\"\n");
  OrientationGenerator og(settings);
  SearchOrCuda search(settings);
  search.Search(og, 0, og.size());
  if (settings.refine_method != SearchOr::Settings::NO_REFINE) search.Refine();
  search.WriteMSF();
  search.WriteChimeraInfo();
  return 0;
}
