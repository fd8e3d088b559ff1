// Registry of kernel families ("modules") that can be run on their own through the C ABI
// (fv3lm_module_run) for parity tests and kernel benchmarks.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "../../include/fv3lm_b200.h"
#include "engine.h"
#include "mosaic.h"

namespace fv3lm {

struct ModuleParams {
  std::map<std::string, double> v;
  const fv3lm_config* cfg = nullptr;
  const std::vector<double>* ak = nullptr;
  const std::vector<double>* bk = nullptr;
  double get(const std::string& k, double dflt) const { auto it = v.find(k); return it == v.end() ? dflt : it->second; }
  int geti(const std::string& k, int dflt) const { return (int)get(k, (double)dflt); }
};

struct ModuleIO {
  std::vector<std::pair<std::string, int>> inputs, outputs;
  int in(Program& P, const std::string& nm, int nk) { int id = P.val(nm, nk, true); inputs.push_back({nm, id}); return id; }
  void out(Program& P, const std::string& nm, int id) { P.vals[id].name = nm; P.vals[id].external = true; outputs.push_back({nm, id}); }
  bool is_input(int id) const { for (auto& kv : inputs) if (kv.second == id) return true; return false; }
};

void build_module(const std::string& name, Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
const char* module_list();

}  // namespace fv3lm
