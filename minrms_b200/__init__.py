"""minrms_b200 -- B200-native MINRMS per-orientation search (hand-written sm_100a CUDA
behind the C ABI of include/minrms_b200.h).  This package is the Python host-side mirror
of the reference's SearchOr interface (bin/minrms/search_or.h:57-130) used by the tests
and the benchmark; the C++ mirror for the reference's CLI is minrms_b200/host/."""
from .search import (Settings, OrientationGenerator, SearchOrCuda, MinrmsError, load_ca_pdb,
                     interpret_num_matches, dp_cells_per_orientation, superimpose_batch)

__all__ = ["Settings", "OrientationGenerator", "SearchOrCuda", "MinrmsError", "load_ca_pdb",
           "interpret_num_matches", "dp_cells_per_orientation", "superimpose_batch"]
