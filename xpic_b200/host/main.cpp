// main.cpp -- `xpic_b200.out <config.json> [-ksp_rtol x ...]`, the drop-in for
// `./run.sh config.json` (run.sh:29, src/main.cpp:9-40) on the path this build covers.
#include <iostream>

#include "xpic_host.h"

int main(int argc, char** argv)
{
  if (argc < 2) {
    std::cerr << "usage: xpic_b200.out <config.json> [-ksp_rtol v] [-ksp_atol v] [-predict_ksp_rtol v] [-correct_ksp_rtol v] [-curl_sign +-1] [-device n] [-precond deg]\n";
    return 2;
  }
  try {
    b200::Simulation simulation;
    for (int i = 2; i + 1 < argc; i += 2) simulation.set_option(argv[i], argv[i + 1]);
    simulation.configure(argv[1]);
    if (simulation.initialize()) return 1;
    if (simulation.calculate()) return 1;
    if (simulation.finalize()) return 1;
  }
  catch (const std::exception& e) {  // src/main.cpp:28-35
    std::cerr << "what(): " << e.what() << "\n";
    return 1;
  }
  return 0;
}
