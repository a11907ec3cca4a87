// xpic_host.h -- C++ host side above the C ABI, mirroring the interface of the reference's
// ecsim::Simulation / ecsimcorr::Simulation / eccapfim::Simulation and interfaces::Particles for this path:
// same member names, same life cycle (initialize / calculate / finalize), same config.json
// schema, same temporal/*.txt output (src/interfaces/simulation.h:21-72,
// src/interfaces/particles.h:11-78, src/diagnostics/energy.cpp, table_diagnostic.h:17-37).
// Errors: every method returns 0 on success (PetscErrorCode convention); configuration errors
// throw std::runtime_error like the reference's builders (src/main.cpp:19-35).
#pragma once
#include <array>
#include <cstdint>
#include <fstream>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/xpic_b200.h"

namespace b200 {

struct Point {  // src/interfaces/point.h:7-35
  double r[3];
  double p[3];
};

struct SortParameters {  // src/interfaces/sort_parameters.h:7-19
  std::string sort_name;
  int Np = 0;
  double n = 0, q = 0, m = 0;
  double px = 0, py = 0, pz = 0;
  double Tx = 0, Ty = 0, Tz = 0;
};

struct Geometry {  // globals of src/constants.h:10-28, set by World::set_geometry (utils/world.cpp:64-85)
  double dx = 0, dy = 0, dz = 0, dt = 0;
  double geom_x = 0, geom_y = 0, geom_z = 0, geom_t = 0;
  int geom_nx = 0, geom_ny = 0, geom_nz = 0, geom_nt = 0;
  int diagnose_period = 1;
};

class Simulation;

class Particles {  // interfaces::Particles + ecsim/ecsimcorr::Particles
public:
  Particles(Simulation& simulation, const SortParameters& parameters, int32_t sid);
  const SortParameters parameters;

  /// src/interfaces/particles.cpp:47-57; points are staged on the host and sent by flush()
  int add_particle(const Point& point, bool* is_added = nullptr);
  int flush();
  /// host mirror of `storage` (cell-major order), refreshed on demand
  int download(std::vector<Point>& out);
  int64_t size();

  // ecsimcorr::Particles scalars (src/impls/ecsimcorr/particles.h:33-38)
  double scalar(int which);

private:
  Simulation& simulation_;
  int32_t sid_;
  std::vector<Point> pending_;
};

/// table_diagnostic.h:17-37 / .cpp:43-59: fixed-width text tables
class Table {
public:
  explicit Table(const std::string& filename, bool append = false);
  void add(int w, std::string title, const std::string& formatted);
  void row(bool with_titles);
  void flush() { file_.flush(); }

private:
  std::ofstream file_;
  std::vector<std::string> titles_, values_;
};

class Simulation {
public:
  Simulation() = default;
  ~Simulation();

  /// reads the reference's config.json schema
  int configure(const std::string& config_path);
  /// options the reference takes from PETSc's database: -ksp_rtol/-ksp_atol/-ksp_max_it (ecsim),
  /// -predict_ksp_*, -correct_ksp_* (ecsimcorr), plus -curl_sign, -device, -precond, and the launch of one
  /// process per GPU: -rank r -nranks N [-comm_file path] (defaults: RANK / WORLD_SIZE / LOCAL_RANK of the environment)
  void set_option(const std::string& key, const std::string& value);

  int initialize();   // interfaces::Simulation::initialize (src/interfaces/simulation.cpp:16-73)
  int calculate();    // :75-96
  int finalize();     // :98-112
  int timestep_implementation(int t);

  /// host mirrors of the named vectors, natural [z][y][x][c] order ("E", "B", "B0", "Ep", "Ec", "currI", "currJe")
  int get_named_vector(const std::string& name, std::vector<double>& out);
  Particles& get_named_particles(const std::string& name);  // throws std::runtime_error if unknown

  Geometry geom;
  int slab_begin() const { return z0_; }  // owned planes [slab_begin, slab_end) of this process
  int slab_end() const { return z0_ + nzl_; }
  std::string scheme_name = "ecsim";
  std::string out_dir = "results";
  std::vector<std::shared_ptr<Particles>> particles_;
  int start = 0;
  xb_ctx* ctx = nullptr;
  int32_t scheme = XB_ECSIM;

private:
  int diagnose_energy(int t);
  int diagnose_convergence(int t);  // eccapfim::ConvergenceHistory
  int diagnose_charge(int t);       // ChargeConservation (ecsimcorr, eccapfim)
  int diagnose_fields(int t);       // FieldView dumps
  int diagnose_momentum(int t);     // MomentumConservation
  int prime_energy();
  int save_backup(int t);           // SimulationBackup::save
  int load_backup(int t);           // SimulationBackup::load
  struct Momentum {  // MomentumGenerator (src/utils/particles_load.h)
    std::string name;
    bool tov = false;
    std::array<double, 3> value{};                                           // PreciseMomentum
    std::array<double, 3> amplitude{}, wave_number{}, box_min{}, box_max{};  // MaxwellCosinePerturbation
  };
  struct Preset {  // a sort with its coordinate and momentum generators
    std::string particles, coordinate;
    Momentum momentum;
    std::array<double, 3> box_min{}, box_max{};       // CoordinateInBox
    std::array<double, 3> center{};                   // CoordinateInCylinder centre / PreciseCoordinate value
    double radius = 0, height = 0;                    // CoordinateInCylinder
  };
  struct Command {  // one entry of "Presets" / "StepPresets" (src/commands)
    std::string name;
    Preset preset;                        // SetParticles, InjectParticles (ionized sort), RemoveParticles (sort name)
    std::string ejected;                  // InjectParticles
    Momentum momentum_e;
    int injection_start = 0, injection_end = 1, tau = 0;
    int64_t per_step = -1;
    int32_t geometry = XB_GEOMETRY_BOX;   // RemoveParticles, FieldsDamping
    std::array<double, 6> p{};
    double coefficient = 0;               // FieldsDamping
    std::string field, field_axpy, setter;  // SetMagneticField
    std::array<double, 3> value{};
    std::vector<std::array<double, 3>> coils;  // {z0, R, I}
  };
  int execute_command(const Command& cmd, int t);
  void draw_coordinate(const Preset& pr, Point& pt);
  void draw_momentum(const Momentum& m, const SortParameters& sp, Point& pt);
  int64_t preset_count(const Preset& pr, const SortParameters& sp) const;
  int diagnose_log(int t);          // LogView (src/diagnostics/log_view.cpp)
  struct View {  // FieldView / DistributionMoment: what is dumped, which region, where
    std::string field, particles, dir, suffix;
    int moment = -1, dof = 3;
    std::array<int, 3> start{}, size{};
  };
  struct VelocityView {  // "Diagnostics": [{"diagnostic": "VelocityDistribution", ...}] (builders/velocity_distribution_builder.cpp)
    std::string particles, projector, dir;
    int32_t projector_id = 0, geometry = XB_GEOMETRY_BOX;
    std::array<double, 6> p{};
    std::array<double, 2> dv{}, vmin{-1.0, -1.0}, vmax{1.0, 1.0};
  };
  std::vector<VelocityView> velocity_views_;
  struct json_ref { const void* p; };  // keeps nlohmann/json out of this header
  void parse_region(const json_ref& info, View& v) const;
  int write_region(const std::string& path, const std::vector<double>& slab, const View& v);  // every rank's part of one dump file
  int log_levels_ = 0;              // bit 0 EachTimestep, bit 1 DiagnosePeriodAvg, bit 2 AllTimestepsSummary
  double log_prev_wall_ = 0, log_period_wall_ = 0, wall_start_ = 0;
  std::array<double, XB_STAGE_COUNT> log_prev_stage_{}, log_period_stage_{};
  std::unique_ptr<std::ofstream> log_each_;
  bool open_z_ = false;             // "da_boundary_z" is not DM_BOUNDARY_PERIODIC
  int da_processors_z_ = -1;        // "mpi": {"da_processors_z"} (utils/configuration.cpp:111-130)
  int rank_ = 0, nranks_ = 1;       // z-slab of this process (-rank / -nranks or RANK / WORLD_SIZE)
  std::string comm_file_;           // where rank 0 leaves the communicator id for the other ranks
  int z0_ = 0, nzl_ = 0;            // owned planes
  std::vector<SortParameters> sorts_;
  std::vector<Command> presets_, step_presets_;
  std::vector<View> moment_views_;  // "Diagnostics": [{"diagnostic": "DistributionMoment", "particles": ..., "moment": ..., "region": ...}]
  std::vector<View> field_views_;   // "Diagnostics": [{"diagnostic": "FieldView", "field": ..., "region": ...}]
  std::unique_ptr<Table> energy_, energy_cons_, convergence_, charge_, momentum_;
  std::vector<std::array<double, 3>> P0_;  // MomentumConservation::P0
  bool charge_header_ = false;
  bool energy_silent_ = false;
  int backup_period_ = 0;  // "SimulationBackup": {"diagnose_period": ...} in steps, 0 = none
  int load_from_ = -1;     // "SimulationBackup": {"load_from": step}
  double E_ = 0, B_ = 0, E0_ = 0, B0_ = 0;
  std::vector<double> K_, K0_, stdK_;
  double stdE_ = 0, stdB_ = 0;
  std::mt19937 gen_;  // src/utils/random_generator.h:20-27, default seed
  std::uniform_real_distribution<double> uni_{0.0, 1.0};
  double rtol_[2] = {1e-7, 1e-7}, atol_[2] = {1e-7, 1e-7};
  int maxit_[2] = {100, 100};
  int curl_sign_ = +1, device_ = 0, precond_ = 6;
  double snes_atol_ = 1e-7, snes_rtol_ = 1e-7, snes_stol_ = 1e-7;  // src/impls/eccapfim/simulation.h:14-19
  int snes_maxit_ = 1000;
};

}  // namespace b200
