// xpic_host.cpp -- see xpic_host.h.
#include "xpic_host.h"

#include <fcntl.h>
#include <unistd.h>

#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <filesystem>
#include <format>
#include <iostream>
#include <stdexcept>
#include <thread>

#include <nlohmann/json.hpp>

namespace b200 {

using json = nlohmann::ordered_json;
constexpr double mec2 = 511.0;  // src/constants.h:30

#define B200_CALL(expr)                                                              \
  do {                                                                               \
    if ((expr) != 0) {                                                               \
      std::cerr << "xpic_b200: " << #expr << " failed: " << xb_last_error() << "\n"; \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

// ---- the unit grammar of Builder::parse_value (src/interfaces/builder.cpp:54-81) ------------------
// a number, one of the geometry names, or "<number> [unit]" with the units of the table below
static double parse_value(const json& value, const Geometry& g)
{
  if (!value.is_string()) return value.get<double>();
  const std::string str = value.get<std::string>();
  const struct { const char* name; double value; } named[] = {
    {"geom_x", g.geom_x}, {"geom_nx", g.geom_x}, {"geom_y", g.geom_y}, {"geom_ny", g.geom_y}, {"geom_z", g.geom_z}, {"geom_nz", g.geom_z}};
  for (const auto& n : named)
    if (str == n.name) return n.value;
  const struct { const char* unit; double scale; } units[] = {
    {" [dx]", g.dx}, {" [dy]", g.dy}, {" [dz]", g.dz}, {" [dt]", g.dt}, {" [c/w_pe]", 1.0}, {" [1/w_pe]", 1.0}};
  for (const auto& u : units) {
    const std::string unit = u.unit;
    if (str.size() > unit.size() && str.compare(str.size() - unit.size(), unit.size(), unit) == 0)
      return std::stod(str.substr(0, str.size() - unit.size())) * u.scale;
  }
  throw std::runtime_error("Unknown string format to convert: " + str);
}

// Builder::parse_vector: three values of the grammar above
static std::array<double, 3> parse_vector(const json& info, const char* key, const Geometry& g)
{
  const json& v = info.at(key);
  if (!v.is_array() || v.size() != 3) throw std::runtime_error(std::string("Vector ") + key + " should have 3 components");
  return {parse_value(v[0], g), parse_value(v[1], g), parse_value(v[2], g)};
}

static int round_step(double s, double ds) { return static_cast<int>(std::round(s / ds)); }

// ---- Table -------------------------------------------------------------------------------------
Table::Table(const std::string& filename, bool append)
{
  if (filename.empty()) return;  // ranks other than 0 evaluate the (collective) diagnostics but do not print them
  std::filesystem::create_directories(std::filesystem::path(filename).parent_path());
  file_.open(filename, append ? std::ios::out | std::ios::app : std::ios::out);
}

void Table::add(int w, std::string title, const std::string& formatted)
{
  titles_.push_back(std::format("{:<{}.{}s}", title, w, w));
  values_.push_back(std::format("{:^{}.{}s}", formatted, w, w));
}

void Table::row(bool with_titles)
{
  auto write = [&](const std::vector<std::string>& c) {
    for (size_t i = 0; i + 1 < c.size(); ++i) file_ << c[i] << "  ";
    std::string last = c.back();
    while (!last.empty() && last.back() == ' ') last.pop_back();
    file_ << last << "\n";
  };
  if (!values_.empty()) {
    if (with_titles) write(titles_);
    write(values_);
  }
  titles_.clear();
  values_.clear();
}

// ---- Particles ---------------------------------------------------------------------------------
Particles::Particles(Simulation& simulation, const SortParameters& parameters, int32_t sid)
  : parameters(parameters), simulation_(simulation), sid_(sid)
{
}

int Particles::add_particle(const Point& point, bool* is_added)
{
  const Geometry& g = simulation_.geom;
  const int vx = (int)std::floor(point.r[0] / g.dx), vy = (int)std::floor(point.r[1] / g.dy), vz = (int)std::floor(point.r[2] / g.dz);
  if (vx < 0 || vx >= g.geom_nx || vy < 0 || vy >= g.geom_ny || vz < simulation_.slab_begin() || vz >= simulation_.slab_end()) return 0;
  pending_.push_back(point);
  if (is_added) *is_added = true;
  return 0;
}

int Particles::flush()
{
  if (pending_.empty()) return 0;
  int64_t added = 0;
  B200_CALL(xb_particles_append(simulation_.ctx, sid_, &pending_[0].r[0], nullptr, (int64_t)pending_.size(), &added));
  pending_.clear();
  return 0;
}

int64_t Particles::size()
{
  int64_t n = 0;
  xb_particles_count(simulation_.ctx, sid_, &n);
  return n;
}

int Particles::download(std::vector<Point>& out)
{
  out.resize((size_t)size());
  int64_t n = 0;
  B200_CALL(xb_particles_download(simulation_.ctx, sid_, out.empty() ? nullptr : &out[0].r[0], nullptr, (int64_t)out.size(), &n));
  return 0;
}

double Particles::scalar(int which)
{
  double v = 0.0;
  xb_scalar(simulation_.ctx, sid_, which, &v);
  return v;
}

// ---- Simulation --------------------------------------------------------------------------------
Simulation::~Simulation() { finalize(); }

void Simulation::set_option(const std::string& key, const std::string& value)
{
  auto tol = [&](const std::string& prefix, int which) {
    if (key == "-" + prefix + "ksp_rtol") { rtol_[which] = std::stod(value); return true; }
    if (key == "-" + prefix + "ksp_atol") { atol_[which] = std::stod(value); return true; }
    if (key == "-" + prefix + "ksp_max_it") { maxit_[which] = std::stoi(value); return true; }
    return false;
  };
  if (tol("", 0) || tol("predict_", 0) || tol("correct_", 1)) return;
  if (key == "-snes_atol") snes_atol_ = std::stod(value);  // SNESSetFromOptions, eccapfim/simulation.cpp:390
  else if (key == "-snes_rtol") snes_rtol_ = std::stod(value);
  else if (key == "-snes_stol") snes_stol_ = std::stod(value);
  else if (key == "-snes_max_it") snes_maxit_ = std::stoi(value);
  else if (key == "-curl_sign") curl_sign_ = std::stoi(value);
  else if (key == "-device") device_ = std::stoi(value);
  else if (key == "-precond") precond_ = std::stoi(value);
  else if (key == "-rank") rank_ = std::stoi(value);
  else if (key == "-nranks") nranks_ = std::stoi(value);
  else if (key == "-comm_file") comm_file_ = value;
  else if (key == "-da_processors_z") da_processors_z_ = std::stoi(value);  // PETSc's own option (DMSetFromOptions)
  else throw std::runtime_error("Unknown option " + key);
}

int Simulation::configure(const std::string& config_path)
{
  std::ifstream f(config_path);
  if (!f) throw std::runtime_error("Cannot open configuration file " + config_path);
  const json cfg = json::parse(f);
  cfg.at("Simulation").get_to(scheme_name);
  if (scheme_name == "ecsim") scheme = XB_ECSIM;
  else if (scheme_name == "ecsimcorr") scheme = XB_ECSIMCORR;
  else if (scheme_name == "eccapfim") scheme = XB_ECCAPFIM;
  else throw std::runtime_error("Unknown simulation is used: " + scheme_name + " (this build covers ecsim, ecsimcorr, eccapfim)");
  if (cfg.contains("OutputDirectory")) cfg.at("OutputDirectory").get_to(out_dir);

  const json& ge = cfg.at("Geometry");  // utils/world.cpp:14-34
  geom.dx = ge.at("dx").get<double>();
  geom.dy = ge.at("dy").get<double>();
  geom.dz = ge.at("dz").get<double>();
  geom.dt = ge.at("dt").get<double>();
  geom.geom_x = parse_value(ge.at("x"), geom);
  geom.geom_y = parse_value(ge.at("y"), geom);
  geom.geom_z = parse_value(ge.at("z"), geom);
  geom.geom_t = parse_value(ge.at("t"), geom);
  geom.geom_nx = round_step(geom.geom_x, geom.dx);
  geom.geom_ny = round_step(geom.geom_y, geom.dy);
  geom.geom_nz = round_step(geom.geom_z, geom.dz);
  geom.geom_nt = round_step(geom.geom_t, geom.dt);
  geom.diagnose_period = std::max(1, round_step(parse_value(ge.at("diagnose_period"), geom), geom.dt));
  for (const char* key : {"da_boundary_x", "da_boundary_y"})
    if (ge.contains(key) && ge.at(key).get<std::string>() != "DM_BOUNDARY_PERIODIC")
      throw std::runtime_error(std::string(key) + ": only DM_BOUNDARY_PERIODIC is covered by this build (z may be DM_BOUNDARY_NONE / GHOSTED)");
  // utils/configuration.cpp:96-104: anything but PERIODIC / GHOSTED means NONE; both have no nodes outside the box
  open_z_ = ge.contains("da_boundary_z") && ge.at("da_boundary_z").get<std::string>() != "DM_BOUNDARY_PERIODIC";

  if (cfg.contains("Particles"))  // src/interfaces/simulation.tpp:12-45
    for (const json& info : cfg.at("Particles")) {
      if (!info.contains("sort_name")) continue;
      SortParameters p;
      info.at("sort_name").get_to(p.sort_name);
      info.at("Np").get_to(p.Np);
      info.at("n").get_to(p.n);
      info.at("q").get_to(p.q);
      info.at("m").get_to(p.m);
      if (info.contains("T")) p.Tx = p.Ty = p.Tz = info.at("T").get<double>();
      else {
        info.at("Tx").get_to(p.Tx);
        info.at("Ty").get_to(p.Ty);
        info.at("Tz").get_to(p.Tz);
      }
      sorts_.push_back(p);
    }
  if (cfg.contains("SimulationBackup") && !cfg.at("SimulationBackup").empty()) {
    // builders/simulation_backup_builder.cpp:38-53 (diagnostic), :63-76 + utils/configuration.cpp:76-86 (restart)
    const json& info = cfg.at("SimulationBackup");
    if (info.contains("diagnose_period")) backup_period_ = std::max(1, round_step(parse_value(info.at("diagnose_period"), geom), geom.dt));
    if (info.contains("load_from") && info.at("load_from").is_number_integer()) load_from_ = info.at("load_from").get<int>();
  }
  if (cfg.contains("Diagnostics"))  // diagnostics/builders/diagnostic_builder.cpp, field_view_builder.cpp:13-52
    for (const json& info : cfg.at("Diagnostics")) {
      const std::string name = info.at("diagnostic").get<std::string>();
      if (name == "FieldView") {  // diagnostics/builders/field_view_builder.cpp:13-50
        View v;
        v.field = info.at("field").get<std::string>();
        v.dof = 3;
        parse_region(json_ref{&info}, v);
        v.dir = out_dir + "/" + v.field + v.suffix;
        field_views_.push_back(v);
      }
      else if (name == "DistributionMoment") {  // diagnostics/builders/distribution_moment_builder.cpp:13-75
        static const std::pair<const char*, std::pair<int, int>> moments[] = {
          {"density", {XB_MOMENT_DENSITY, 1}}, {"current", {XB_MOMENT_CURRENT, 3}}, {"momentum_flux", {XB_MOMENT_MOMENTUM_FLUX, 6}},
          {"momentum_flux_cyl", {XB_MOMENT_MOMENTUM_FLUX_CYL, 6}}, {"momentum_flux_diag", {XB_MOMENT_MOMENTUM_FLUX_DIAG, 3}},
          {"momentum_flux_diag_cyl", {XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL, 3}}};
        View v;
        v.particles = info.at("particles").get<std::string>();
        v.field = info.at("moment").get<std::string>();
        v.moment = -1;
        for (const auto& m : moments)
          if (v.field == m.first) {
            v.moment = m.second.first;
            v.dof = m.second.second;
          }
        if (v.moment < 0) throw std::runtime_error("Unknown moment name " + v.field + " for particles " + v.particles);
        parse_region(json_ref{&info}, v);
        v.dir = out_dir + "/" + v.particles + "/" + v.field + v.suffix;
        moment_views_.push_back(v);
      }
      else if (name == "VelocityDistribution") {  // diagnostics/builders/velocity_distribution_builder.cpp:13-112
        static const std::pair<const char*, int> projectors[] = {{"vx_vy", XB_PROJECTOR_VX_VY}, {"vz_vxy", XB_PROJECTOR_VZ_VXY}, {"vr_vphi", XB_PROJECTOR_VR_VPHI}};
        VelocityView v;
        info.at("particles").get_to(v.particles);
        info.at("projector").get_to(v.projector);
        v.projector_id = -1;
        for (const auto& pr : projectors)
          if (v.projector == pr.first) v.projector_id = pr.second;
        if (v.projector_id < 0) throw std::runtime_error("Unkown projector name " + v.projector);  // the reference's message, velocity_distribution.cpp:204
        const json& ge = info.at("geometry");
        const std::string gname = ge.at("name").get<std::string>();
        if (gname == "BoxGeometry") {  // Builder::load_geometry, interfaces/builder.cpp:83-94
          std::array<double, 3> lo = {0.0, 0.0, 0.0}, hi = {geom.geom_x, geom.geom_y, geom.geom_z};
          if (ge.contains("min")) lo = parse_vector(ge, "min", geom);
          if (ge.contains("max")) hi = parse_vector(ge, "max", geom);
          v.geometry = XB_GEOMETRY_BOX;
          v.p = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
        }
        else if (gname == "CylinderGeometry") {  // :96-111
          std::array<double, 3> center = {0.5 * geom.geom_x, 0.5 * geom.geom_y, 0.5 * geom.geom_z};
          double radius = 0.5 * std::min(geom.geom_x, geom.geom_y), height = geom.geom_z;
          if (ge.contains("center")) center = parse_vector(ge, "center", geom);
          if (ge.contains("radius")) ge.at("radius").get_to(radius);
          if (ge.contains("height")) ge.at("height").get_to(height);
          v.geometry = XB_GEOMETRY_CYLINDER;
          v.p = {center[0], center[1], center[2], radius, height, 0.0};
        }
        else
          throw std::runtime_error("Unknown geometry name " + gname);
        info.at("dv").at(0).get_to(v.dv[0]);
        info.at("dv").at(1).get_to(v.dv[1]);
        if (info.contains("vmax")) v.vmax = {info.at("vmax").at(0).get<double>(), info.at("vmax").at(1).get<double>()};
        if (info.contains("vmin")) v.vmin = {info.at("vmin").at(0).get<double>(), info.at("vmin").at(1).get<double>()};
        v.dir = out_dir + "/" + v.particles + "/" + v.projector;
        velocity_views_.push_back(v);
      }
      else if (name == "LogView") {  // diagnostics/builders/log_view_builder.cpp: one of three levels
        const std::string level = info.at("level").get<std::string>();
        if (level == "EachTimestep") log_levels_ |= 1;
        else if (level == "DiagnosePeriodAvg") log_levels_ |= 2;
        else if (level == "AllTimestepsSummary") log_levels_ |= 4;
        else throw std::runtime_error("Unknown LogView level " + level);
      }
      else
        std::cout << "  diagnostic " << name << " is not covered by this build, skipped\n";
    }
  // ParticlesBuilder::load_coordinate / load_momentum (commands/builders/particles_builder.cpp:9-70)
  auto load_coordinate = [&](const json& co, Preset& pr) {
    pr.coordinate = co.at("name").get<std::string>();
    pr.box_min = {0.0, 0.0, 0.0};
    pr.box_max = {geom.geom_x, geom.geom_y, geom.geom_z};
    if (pr.coordinate == "CoordinateInBox") {
      if (co.contains("min")) pr.box_min = parse_vector(co, "min", geom);
      if (co.contains("max")) pr.box_max = parse_vector(co, "max", geom);
    }
    else if (pr.coordinate == "CoordinateInCylinder") {  // builder.cpp:96-111
      pr.center = {0.5 * geom.geom_x, 0.5 * geom.geom_y, 0.5 * geom.geom_z};
      pr.radius = 0.5 * std::min(geom.geom_x, geom.geom_y);
      pr.height = geom.geom_z;
      if (co.contains("center")) pr.center = parse_vector(co, "center", geom);
      if (co.contains("radius")) co.at("radius").get_to(pr.radius);
      if (co.contains("height")) co.at("height").get_to(pr.height);
    }
    else if (pr.coordinate == "PreciseCoordinate")
      pr.center = parse_vector(co, "value", geom);
    else
      throw std::runtime_error("Unknown coordinate generator name " + pr.coordinate);
  };
  auto load_momentum = [&](const json& mo, Momentum& m) {
    m.name = mo.at("name").get<std::string>();
    if (m.name == "MaxwellianMomentum") {
      if (mo.contains("tov")) mo.at("tov").get_to(m.tov);
    }
    else if (m.name == "PreciseMomentum")
      m.value = parse_vector(mo, "value", geom);
    else if (m.name == "MaxwellCosinePerturbation") {
      m.box_min = {0.0, 0.0, 0.0};
      m.box_max = {geom.geom_x, geom.geom_y, geom.geom_z};
      if (mo.contains("min")) m.box_min = parse_vector(mo, "min", geom);
      if (mo.contains("max")) m.box_max = parse_vector(mo, "max", geom);
      m.amplitude = parse_vector(mo, "amplitude", geom);
      m.wave_number = parse_vector(mo, "wave_number", geom);
    }
    else
      throw std::runtime_error("Unknown coordinate generator name " + m.name);  // the reference's message, particles_builder.cpp:67
  };
  // Builder::load_geometry (interfaces/builder.cpp:83-111) into the C ABI's six numbers
  auto load_geometry = [&](const json& ge, Command& cmd) {
    const std::string name = ge.at("name").get<std::string>();
    if (name == "BoxGeometry") {
      std::array<double, 3> lo = {0.0, 0.0, 0.0}, hi = {geom.geom_x, geom.geom_y, geom.geom_z};
      if (ge.contains("min")) lo = parse_vector(ge, "min", geom);
      if (ge.contains("max")) hi = parse_vector(ge, "max", geom);
      cmd.geometry = XB_GEOMETRY_BOX;
      cmd.p = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    }
    else if (name == "CylinderGeometry") {
      std::array<double, 3> center = {0.5 * geom.geom_x, 0.5 * geom.geom_y, 0.5 * geom.geom_z};
      double radius = 0.5 * std::min(geom.geom_x, geom.geom_y), height = geom.geom_z;
      if (ge.contains("center")) center = parse_vector(ge, "center", geom);
      if (ge.contains("radius")) ge.at("radius").get_to(radius);
      if (ge.contains("height")) ge.at("height").get_to(height);
      cmd.geometry = XB_GEOMETRY_CYLINDER;
      cmd.p = {center[0], center[1], center[2], radius, height, 0.0};
    }
    else
      throw std::runtime_error("Unknown geometry name " + name);
  };
  // commands/builders/command_builder.cpp:42-59: "Presets" run once before the first step, "StepPresets" before every step
  for (const char* list : {"Presets", "StepPresets"}) {
    if (!cfg.contains(list)) continue;
    const bool every_step = std::string(list) == "StepPresets";
    for (const json& info : cfg.at(list)) {
      const std::string command = info.at("command").get<std::string>();
      Command cmd;
      cmd.name = command;
      if (command == "SetParticles") {
        info.at("particles").get_to(cmd.preset.particles);
        load_coordinate(info.at("coordinate"), cmd.preset);
        load_momentum(info.at("momentum"), cmd.preset.momentum);
      }
      else if (command == "InjectParticles") {  // commands/builders/inject_particles_builder.cpp:9-68
        info.at("ionized").get_to(cmd.preset.particles);
        info.at("ejected").get_to(cmd.ejected);
        if (info.contains("injection_start")) cmd.injection_start = round_step(info.at("injection_start").get<double>(), geom.dt);
        if (info.contains("injection_end")) {
          const json& v = info.at("injection_end");
          if (v.is_string()) {
            if (v.get<std::string>() == "geom_t") cmd.injection_end = geom.geom_nt;
          }
          else
            cmd.injection_end = round_step(v.get<double>(), geom.dt);
        }
        load_coordinate(info.at("coordinate"), cmd.preset);
        load_momentum(info.at("momentum_i"), cmd.preset.momentum);
        load_momentum(info.at("momentum_e"), cmd.momentum_e);
        if (info.contains("tau")) cmd.tau = round_step(info.at("tau").get<double>(), geom.dt);
        if (info.contains("per_step_particles_num")) cmd.per_step = info.at("per_step_particles_num").get<int64_t>();
      }
      else if (command == "RemoveParticles") {  // commands/builders/remove_particles_builder.cpp
        info.at("particles").get_to(cmd.preset.particles);
        load_geometry(info.at("geometry"), cmd);
      }
      else if (command == "FieldsDamping") {  // commands/builders/fields_damping_builder.cpp
        for (const char* key : {"E", "B", "B0"})
          if (info.at(key).get<std::string>() != key) throw std::runtime_error("FieldsDamping: the vectors must be E, B, B0");
        info.at("damping_coefficient").get_to(cmd.coefficient);
        load_geometry(info.at("geometry"), cmd);
      }
      else if (command == "SetMagneticField") {  // commands/builders/set_magnetic_field_builder.cpp
        info.at("field").get_to(cmd.field);
        if (info.contains("field_axpy")) info.at("field_axpy").get_to(cmd.field_axpy);
        const json& setter = info.at("setter");
        cmd.setter = setter.at("name").get<std::string>();
        if (cmd.setter == "SetUniformField")
          cmd.value = parse_vector(setter, "value", geom);
        else if (cmd.setter == "SetCoilsField")
          for (const json& coil : setter.at("coils")) cmd.coils.push_back({coil.at("z0").get<double>(), coil.at("R").get<double>(), coil.at("I").get<double>()});
        else
          throw std::runtime_error("Unknown setter name " + cmd.setter);
      }
      else
        throw std::runtime_error("Preset " + command + " is not covered by this build");
      (every_step ? step_presets_ : presets_).push_back(cmd);
    }
  }
  if (cfg.contains("mpi")) {  // utils/configuration.cpp:111-130: only the z split exists in a slab layout
    const json& mpi = cfg.at("mpi");
    for (const char* key : {"da_processors_x", "da_processors_y"})
      if (mpi.contains(key) && mpi.at(key).get<int>() > 1)
        throw std::runtime_error(std::string(key) + " > 1: this build decomposes along z only (da_processors_z)");
    if (mpi.contains("da_processors_z")) da_processors_z_ = mpi.at("da_processors_z").get<int>();
  }
  return 0;
}

// "region" of a FieldView / DistributionMoment (field_view_builder.cpp:52-147): a 3D box or a one-cell-thick 2D plane
void Simulation::parse_region(const json_ref& info_ref, View& v) const
{
  const json& info = *static_cast<const json*>(info_ref.p);
  v.start = {0, 0, 0};
  v.size = {geom.geom_nx, geom.geom_ny, geom.geom_nz};
  v.suffix.clear();
  if (!info.contains("region")) return;
  const json& r = info.at("region");
  const std::string type = r.contains("type") ? r.at("type").get<std::string>() : std::string("3D");
  if (type != "3D" && type != "2D") throw std::runtime_error("Incorrect type is used for " + v.field + " .");
  std::array<double, 3> start = {0.0, 0.0, 0.0}, size = {geom.geom_x, geom.geom_y, geom.geom_z};
  if (r.contains("start")) start = parse_vector(r, "start", geom);
  if (r.contains("size")) size = parse_vector(r, "size", geom);
  const double d[3] = {geom.dx, geom.dy, geom.dz};
  if (type == "2D") {
    const std::string plane = r.at("plane").get<std::string>();
    const int axis = plane == "X" ? 0 : (plane == "Y" ? 1 : (plane == "Z" ? 2 : -1));
    if (axis < 0) throw std::runtime_error("Unknown plane " + plane);
    double position = 0.5 * (axis == 0 ? geom.geom_x : (axis == 1 ? geom.geom_y : geom.geom_z));
    if (r.contains("position")) r.at("position").get_to(position);
    start[axis] = position;
    size[axis] = d[axis];
    v.suffix = std::format("_plane{}_{:04d}", plane, (int)std::floor(position / d[axis]));
  }
  const int n[3] = {geom.geom_nx, geom.geom_ny, geom.geom_nz};
  for (int a = 0; a < 3; ++a) {
    v.start[a] = (int)std::floor(start[a] / d[a]);
    v.size[a] = (int)std::floor(size[a] / d[a]);
    if (v.start[a] < 0 || v.start[a] + v.size[a] > n[a]) throw std::runtime_error("Region is not in global boundaries for " + v.field + " diagnostic.");
    if (v.size[a] <= 0) throw std::runtime_error("Sizes are invalid for " + v.field + " diagnostic.");
  }
}

Particles& Simulation::get_named_particles(const std::string& name)
{
  for (auto& s : particles_)
    if (s->parameters.sort_name == name) return *s;
  throw std::runtime_error("No particles with name " + name);  // src/interfaces/simulation.cpp:145-155
}

int Simulation::get_named_vector(const std::string& name, std::vector<double>& out)
{
  static const std::vector<std::pair<std::string, int>> names = {{"E", XB_E}, {"B", XB_B}, {"B0", XB_B0}, {"Ep", XB_EP}, {"Ec", XB_EC},
                                                                 {"currI", XB_CURRI}, {"currJe", XB_CURRJE}};
  for (auto& [n, id] : names)
    if (n == name) {
      out.resize((size_t)3 * geom.geom_nx * geom.geom_ny * nzl_);  // this rank's z-slab (the whole box on one GPU)
      B200_CALL(xb_field_download(ctx, id, 0, out.data()));
      return 0;
    }
  throw std::runtime_error("Unknown vector name " + name);
}

int Simulation::initialize()
{
  xb_grid g{};
  g.n[0] = geom.geom_nx; g.n[1] = geom.geom_ny; g.n[2] = geom.geom_nz;
  g.d[0] = geom.dx; g.d[1] = geom.dy; g.d[2] = geom.dz;
  g.dt = geom.dt;
  g.curl_sign = curl_sign_;
  g.device = device_;
  // one process per GPU (the reference: one MPI rank per DMDA box, -da_processors_z N); the launcher passes
  // -rank / -nranks or sets RANK / WORLD_SIZE / LOCAL_RANK (torchrun, mpirun wrappers)
  auto env_int = [](const char* name, int fallback) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : fallback;
  };
  if (nranks_ == 1 && rank_ == 0) {
    nranks_ = env_int("WORLD_SIZE", 1);
    rank_ = env_int("RANK", 0);
    if (nranks_ > 1 && device_ == 0) device_ = env_int("LOCAL_RANK", rank_);
  }
  if (da_processors_z_ > 0 && da_processors_z_ != nranks_)
    throw std::runtime_error(std::format("da_processors_z = {} but {} process(es) were started: launch one process per z-slab (-rank r -nranks N)",
                                         da_processors_z_, nranks_));
  g.rank = rank_;
  g.nranks = nranks_;
  g.track_ids = 0;
  g.boundary[2] = open_z_ ? XB_BOUNDARY_OPEN : XB_BOUNDARY_PERIODIC;
  {  // the library's split (DMDA's: the first nz % N slabs are one plane thicker)
    const int base = geom.geom_nz / nranks_, rem = geom.geom_nz % nranks_;
    nzl_ = base + (rank_ < rem ? 1 : 0);
    z0_ = rank_ * base + std::min(rank_, rem);
  }
  unsigned char uid[128];
  const void* uid_ptr = nullptr;
  if (nranks_ > 1) {
    // MPI_Init's role: rank 0 creates the NCCL id and leaves it in a file the other ranks wait for
    if (comm_file_.empty()) comm_file_ = out_dir + "/.xpic_b200_comm_id";
    if (rank_ == 0) {
      B200_CALL(xb_comm_unique_id(uid));
      std::filesystem::create_directories(std::filesystem::path(comm_file_).parent_path());
      std::ofstream(comm_file_ + ".tmp", std::ios::binary).write(reinterpret_cast<const char*>(uid), sizeof(uid));
      std::filesystem::rename(comm_file_ + ".tmp", comm_file_);
    }
    else {
      for (int tries = 0; !std::filesystem::exists(comm_file_) || std::filesystem::file_size(comm_file_) != sizeof(uid); ++tries) {
        if (tries > 6000) throw std::runtime_error("rank " + std::to_string(rank_) + ": no communicator id in " + comm_file_);
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
      std::ifstream(comm_file_, std::ios::binary).read(reinterpret_cast<char*>(uid), sizeof(uid));
    }
    uid_ptr = uid;
  }
  if (rank_ != 0) std::cout.setstate(std::ios_base::failbit);  // PetscPrintf(PETSC_COMM_WORLD, ...): rank 0 talks
  B200_CALL(xb_create(&g, uid_ptr, &ctx));
  if (nranks_ > 1 && rank_ == 0) std::filesystem::remove(comm_file_);  // every rank has joined the communicator
  for (int w = 0; w < 2; ++w) B200_CALL(xb_solver_set(ctx, w, rtol_[w], atol_[w], maxit_[w], 30, precond_));
  // SNESSetTolerances + the Crank-Nicolson tolerance 0.5 * atol (eccapfim/simulation.cpp:384, particles.cpp:99-101)
  B200_CALL(xb_nonlinear_set(ctx, snes_atol_, snes_rtol_, snes_stol_, snes_maxit_, 10, 12, 0.5 * snes_atol_, 30));

  const int64_t ncells = (int64_t)geom.geom_nx * geom.geom_ny * nzl_;
  for (const auto& p : sorts_) {
    int32_t sid = 0;
    B200_CALL(xb_species_add(ctx, p.q, p.m, p.n, p.Np, (int64_t)(1.5 * ncells * p.Np) + 4096, &sid));
    particles_.push_back(std::make_shared<Particles>(*this, p, sid));
  }

  // presets: SetParticles::execute (src/commands/set_particles.cpp:19-43) with the generators of
  // src/utils/particles_load.cpp:11-18,52-76 and the count of particles_builder.cpp:17,26
  const bool restart = load_from_ >= 0;  // "Other preset commands would be dropped" (simulation_backup_builder.cpp:70)
  if (restart) {
    if (load_backup(load_from_)) return 1;
    start = load_from_;
  }
  if (!restart)
    for (const Command& cmd : presets_)
      if (execute_command(cmd, start)) return 1;

  // after a restart the tables of the backup continue (rows up to `start` are already there)
  energy_ = std::make_unique<Table>(rank_ == 0 ? out_dir + "/temporal/energy.txt" : std::string(), restart);
  energy_cons_ = std::make_unique<Table>(rank_ == 0 ? out_dir + "/temporal/energy_conservation.txt" : std::string(), restart);
  K_.assign(particles_.size(), 0.0);
  K0_ = stdK_ = K_;
  if (scheme == XB_ECCAPFIM) {  // eccapfim/simulation.cpp:28
    convergence_ = std::make_unique<Table>(rank_ == 0 ? out_dir + "/temporal/convergence_history.txt" : std::string(), restart);
    if (!restart && diagnose_convergence(start)) return 1;
  }
  if (scheme != XB_ECSIM) {  // ChargeConservation: interfaces/simulation.cpp:32-38 (J), ecsimcorr/simulation.cpp:103-110 (currJe)
    charge_ = std::make_unique<Table>(rank_ == 0 ? out_dir + "/temporal/charge_conservation.txt" : std::string(), restart);
    charge_header_ = restart;
    for (size_t i = 0; i < particles_.size(); ++i) B200_CALL(xb_charge_density(ctx, (int32_t)i, nullptr));  // ChargeConservation::initialize
    if (!restart && diagnose_charge(start)) return 1;
  }
  // MomentumConservation (interfaces/simulation.cpp:54-56): initialize() stores P at t = 0
  momentum_ = std::make_unique<Table>(rank_ == 0 ? out_dir + "/temporal/momentum_conservation.txt" : std::string(), restart);
  P0_.assign(particles_.size(), {0.0, 0.0, 0.0});
  for (size_t i = 0; i < particles_.size(); ++i) {
    double o[6];
    B200_CALL(xb_momentum(ctx, (int32_t)i, o));
    P0_[i] = {o[0], o[1], o[2]};
  }
  if (restart) {
    // prime the "previous step" values of the table diagnostics without writing the row of t = start again
    return prime_energy();
  }
  if (diagnose_momentum(start)) return 1;
  if (diagnose_fields(start)) return 1;
  // the backup is the last diagnostic of the reference's list: its copy of temporal/ already holds row `start`
  if (diagnose_energy(start)) return 1;
  if (backup_period_ > 0 && save_backup(start)) return 1;
  return 0;
}

// ---- commands (src/commands) ------------------------------------------------------------------------------
// Coordinates and momenta come from the reference's default-seeded mt19937 stream in its draw order
// (src/utils/particles_load.cpp); every rank draws the whole stream and keeps its own particles.
void Simulation::draw_coordinate(const Preset& pr, Point& pt)
{
  auto r01 = [&]() { return uni_(gen_); };
  if (pr.coordinate == "CoordinateInBox") {
    for (int c = 0; c < 3; ++c) pt.r[c] = pr.box_min[c] + r01() * (pr.box_max[c] - pr.box_min[c]);
  }
  else if (pr.coordinate == "CoordinateInCylinder") {
    const double r = pr.radius * std::sqrt(r01());
    const double phi = 2.0 * M_PI * r01();
    pt.r[0] = pr.center[0] + r * std::cos(phi);
    pt.r[1] = pr.center[1] + r * std::sin(phi);
    pt.r[2] = pr.center[2] + pr.height * (r01() - 0.5);
  }
  else {
    for (int c = 0; c < 3; ++c) pt.r[c] = pr.center[c];
  }
}

void Simulation::draw_momentum(const Momentum& m, const SortParameters& sp, Point& pt)
{
  auto r01 = [&]() { return uni_(gen_); };
  if (m.name == "PreciseMomentum") {
    for (int c = 0; c < 3; ++c) pt.p[c] = m.value[c];
    return;
  }
  auto tm = [&](double T) { return std::sqrt(-2.0 * (T * sp.m / mec2) * std::log(r01())); };
  const double T[3] = {sp.Tx, sp.Ty, sp.Tz}, p0[3] = {sp.px, sp.py, sp.pz};
  const bool cosine = m.name == "MaxwellCosinePerturbation";  // particles_load.cpp:78-104
  for (int c = 0; c < 3; ++c) {
    const double sn = std::sin(2.0 * M_PI * r01());  // the sine factor is drawn first
    pt.p[c] = (cosine ? 0.0 : p0[c]) + sn * tm(T[c]);
  }
  if (m.tov || cosine) {
    const double den = std::sqrt(sp.m * sp.m + (pt.p[0] * pt.p[0] + pt.p[1] * pt.p[1] + pt.p[2] * pt.p[2]));
    for (double& v : pt.p) v /= den;
  }
  if (cosine)
    for (int c = 0; c < 3; ++c) {
      const double v0 = m.amplitude[c] * std::sqrt(T[c] / (sp.m * mec2));
      pt.p[c] += v0 * std::cos(2.0 * M_PI * m.wave_number[c] * pt.r[c] / (m.box_max[c] - m.box_min[c]));
    }
}

// number_of_particles of ParticlesBuilder::load_coordinate (particles_builder.cpp:16-37), truncated like its PetscInt
int64_t Simulation::preset_count(const Preset& pr, const SortParameters& sp) const
{
  const double frac = sp.Np / (geom.dx * geom.dy * geom.dz);
  if (pr.coordinate == "CoordinateInBox")
    return (int64_t)(((pr.box_max[0] - pr.box_min[0]) * (pr.box_max[1] - pr.box_min[1]) * (pr.box_max[2] - pr.box_min[2])) * frac);
  if (pr.coordinate == "CoordinateInCylinder") return (int64_t)(M_PI * (pr.radius * pr.radius) * pr.height * frac);
  return sp.Np;
}

int Simulation::execute_command(const Command& cmd, int t)
{
  if (cmd.name == "SetParticles") {  // SetParticles::execute (src/commands/set_particles.cpp:19-43)
    Particles& sort = get_named_particles(cmd.preset.particles);
    const int64_t count = preset_count(cmd.preset, sort.parameters);
    for (int64_t i = 0; i < count; ++i) {
      Point pt;
      draw_coordinate(cmd.preset, pt);
      draw_momentum(cmd.preset.momentum, sort.parameters, pt);
      sort.add_particle(pt);
    }
    return sort.flush();
  }
  if (cmd.name == "InjectParticles") {  // InjectParticles::execute (src/commands/inject_particles.cpp:27-67)
    if (t < cmd.injection_start || t > cmd.injection_end) return 0;
    Particles& ionized = get_named_particles(cmd.preset.particles);
    Particles& ejected = get_named_particles(cmd.ejected);
    int64_t per_step = cmd.per_step;
    if (per_step < 0) {
      const int tau = cmd.tau > 0 ? cmd.tau : cmd.injection_end - cmd.injection_start;
      per_step = preset_count(cmd.preset, ionized.parameters) / std::max(tau, 1);
    }
    for (int64_t i = 0; i < per_step; ++i) {
      Point pi, pe;
      draw_coordinate(cmd.preset, pi);
      for (int c = 0; c < 3; ++c) pe.r[c] = pi.r[c];
      draw_momentum(cmd.preset.momentum, ionized.parameters, pi);
      draw_momentum(cmd.momentum_e, ejected.parameters, pe);
      ionized.add_particle(pi);
      ejected.add_particle(pe);
    }
    if (ionized.flush()) return 1;
    return ejected.flush();
  }
  if (cmd.name == "RemoveParticles") {  // RemoveParticles::execute (src/commands/remove_particles.cpp:11-45) on the device
    int32_t sid = -1;
    for (size_t i = 0; i < particles_.size(); ++i)
      if (particles_[i]->parameters.sort_name == cmd.preset.particles) sid = (int32_t)i;
    if (sid < 0) throw std::runtime_error("No particles with name " + cmd.preset.particles);
    double out[2] = {0.0, 0.0};
    B200_CALL(xb_particles_remove(ctx, sid, cmd.geometry, cmd.p.data(), out));
    std::cout << std::format("  Particles have been removed from \"{}\": {} particles, energy: {:6.4e}", cmd.preset.particles, (int64_t)out[0], out[1]) << "\n";
    return 0;
  }
  if (cmd.name == "FieldsDamping") {  // FieldsDamping::execute (src/commands/fields_damping.cpp:16-31) on the device
    double taken = 0.0;
    B200_CALL(xb_fields_damping(ctx, cmd.geometry, cmd.p.data(), cmd.coefficient, &taken));
    std::cout << std::format("  Fields are damped, additional energy runoff: {:6.4e}", taken) << "\n";
    return 0;
  }
  if (cmd.name == "SetMagneticField") {  // SetMagneticField::execute (src/commands/set_magnetic_field.cpp:12-19)
    const size_t nloc = (size_t)3 * geom.geom_nx * geom.geom_ny * nzl_;
    std::vector<double> f;
    if (get_named_vector(cmd.field, f)) return 1;
    if (cmd.setter == "SetUniformField") {  // VecStrideSet: the value replaces what was there (:27-35)
      for (size_t i = 0; i < nloc; ++i) f[i] = cmd.value[i % 3];
    }
    else {  // SetCoilsField::operator() (:47-91): a sum of current loops on the axis of the box, N = 2000 point quadrature
      constexpr int N = 2000;
      constexpr double hp = 2 * M_PI / N, tol = 1e-10;
      std::vector<double> cosv(N);
      for (int i = 0; i < N; ++i) cosv[i] = std::cos(i * hp);
      auto integ = [&](double z, double r, double R, bool radial) {
        double integral = 0.0;
        for (int i = 0; i < N; ++i) {
          double den = z * z + R * R + r * r - 2.0 * R * r * cosv[i];
          if (std::abs(den) < tol) den = tol;
          integral += (radial ? cosv[i] : (R - r * cosv[i])) / (den * std::sqrt(den));
        }
        return hp * integral;
      };
      auto Br = [&](double z, double r) {
        double v = 0.0;
        for (auto& co : cmd.coils) v += co[2] * co[1] * (z - co[0]) * integ(z - co[0], r, co[1], true);
        return v;
      };
      auto Bz = [&](double z, double r) {
        double v = 0.0;
        for (auto& co : cmd.coils) v += co[2] * co[1] * integ(z - co[0], r, co[1], false);
        return v;
      };
      const double cx = 0.5 * geom.geom_x, cy = 0.5 * geom.geom_y;
      for (int zl = 0; zl < nzl_; ++zl)
        for (int y = 0; y < geom.geom_ny; ++y)
          for (int x = 0; x < geom.geom_nx; ++x) {
            const int z = z0_ + zl;
            double* a = &f[(((size_t)zl * geom.geom_ny + y) * geom.geom_nx + x) * 3];
            double sx = x * geom.dx - cx, sy = (y + 0.5) * geom.dy - cy, sz = (z + 0.5) * geom.dz, r = std::hypot(sx, sy);
            a[0] += Br(sz, r) * sx / r;
            sy = y * geom.dy - cy;
            sx = (x + 0.5) * geom.dx - cx;
            r = std::hypot(sx, sy);
            a[1] += Br(sz, r) * sy / r;
            sz = z * geom.dz;
            sy = (y + 0.5) * geom.dy - cy;
            r = std::hypot(sx, sy);
            a[2] += Bz(sz, r);
          }
    }
    static const std::vector<std::pair<std::string, int>> names = {{"E", XB_E}, {"B", XB_B}, {"B0", XB_B0}};
    auto id_of = [&](const std::string& n) {
      for (auto& [name, id] : names)
        if (name == n) return id;
      throw std::runtime_error("Unknown vector name " + n);
    };
    B200_CALL(xb_field_upload(ctx, id_of(cmd.field), 0, f.data()));
    if (!cmd.field_axpy.empty()) {  // VecAXPY(B, 1.0, B0)
      std::vector<double> b;
      if (get_named_vector(cmd.field_axpy, b)) return 1;
      for (size_t i = 0; i < nloc; ++i) b[i] += f[i];
      B200_CALL(xb_field_upload(ctx, id_of(cmd.field_axpy), 0, b.data()));
    }
    return 0;
  }
  throw std::runtime_error("Preset " + cmd.name + " is not covered by this build");
}

// FieldView::diagnose (src/diagnostics/field_view.cpp:98-118): float32 image of the whole vector in
// natural [z][y][x][c] order (MPIBinaryFile::write_floats, utils/mpi_binary_file.cpp:98-106), named by
// Diagnostic::format_time (interfaces/diagnostic.cpp:21-25)
int Simulation::diagnose_fields(int t)
{
  if (t % geom.diagnose_period != 0) return 0;
  const int width = (int)std::to_string(geom.geom_nt).size();
  std::vector<double> f;
  for (const View& v : moment_views_) {  // DistributionMoment::diagnose (distribution_moment.cpp:112-122)
    int32_t sid = -1;
    for (size_t i = 0; i < particles_.size(); ++i)
      if (particles_[i]->parameters.sort_name == v.particles) sid = (int32_t)i;
    if (sid < 0) throw std::runtime_error("No particles with name " + v.particles);
    f.resize((size_t)v.dof * geom.geom_nx * geom.geom_ny * nzl_);
    const int32_t st[3] = {v.start[0], v.start[1], v.start[2]}, sz[3] = {v.size[0], v.size[1], v.size[2]};
    B200_CALL(xb_distribution_moment_region(ctx, sid, v.moment, st, sz, f.data()));
    std::filesystem::create_directories(v.dir);
    if (write_region(v.dir + "/" + std::format("{:0{}d}", t, width), f, v)) return 1;
  }
  for (const VelocityView& v : velocity_views_) {  // VelocityDistribution::collect + FieldView::diagnose: one float32 [vy][vx] image
    int32_t sid = -1;
    for (size_t i = 0; i < particles_.size(); ++i)
      if (particles_[i]->parameters.sort_name == v.particles) sid = (int32_t)i;
    if (sid < 0) throw std::runtime_error("No particles with name " + v.particles);
    int32_t vstart = 0, vsize = 0;
    B200_CALL(xb_velocity_distribution_size(v.dv.data(), v.vmin.data(), v.vmax.data(), &vstart, &vsize));
    f.assign((size_t)vsize * vsize, 0.0);
    B200_CALL(xb_velocity_distribution(ctx, sid, v.projector_id, v.geometry, v.p.data(), v.dv.data(), v.vmin.data(), v.vmax.data(), f.data()));
    if (rank_ == 0) {  // every rank holds the sum
      std::filesystem::create_directories(v.dir);
      std::vector<float> img(f.begin(), f.end());
      std::ofstream out(v.dir + "/" + std::format("{:0{}d}", t, width), std::ios::binary | std::ios::trunc);
      out.write(reinterpret_cast<const char*>(img.data()), (std::streamsize)(img.size() * sizeof(float)));
      if (!out) throw std::runtime_error("short write into " + v.dir);
    }
  }
  for (const View& v : field_views_) {
    if (get_named_vector(v.field, f)) return 1;
    std::filesystem::create_directories(v.dir);
    if (write_region(v.dir + "/" + std::format("{:0{}d}", t, width), f, v)) return 1;
  }
  return 0;
}

// One dump file for all ranks, as MPIBinaryFile writes it through MPI subarray views (field_view.cpp:59-95): the
// float32 image of the region in natural [z][y][x][component] order.  `slab` is this rank's z-slab of the whole box;
// the rows of the region that lie in it are written at their offsets (one write per plane when the region spans x and y).
int Simulation::write_region(const std::string& path, const std::vector<double>& slab, const View& v)
{
  const int nx = geom.geom_nx, ny = geom.geom_ny;
  const int zlo = std::max(v.start[2], z0_), zhi = std::min(v.start[2] + v.size[2], z0_ + nzl_);
  const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT, 0644);
  if (fd < 0) throw std::runtime_error("cannot open " + path + " for writing");
  const size_t row = (size_t)v.size[0] * v.dof;
  std::vector<float> buf(row * v.size[1]);
  bool ok = true;
  for (int z = zlo; z < zhi && ok; ++z) {
    for (int y = 0; y < v.size[1]; ++y) {
      const double* src = slab.data() + ((((size_t)(z - z0_) * ny + (v.start[1] + y)) * nx + v.start[0]) * v.dof);
      for (size_t i = 0; i < row; ++i) buf[(size_t)y * row + i] = (float)src[i];
    }
    const off_t at = (off_t)sizeof(float) * row * v.size[1] * (off_t)(z - v.start[2]);
    const char* data = reinterpret_cast<const char*>(buf.data());
    size_t left = buf.size() * sizeof(float), done = 0;
    while (ok && left > 0) {
      const ssize_t w = ::pwrite(fd, data + done, left, at + (off_t)done);
      ok = w > 0;
      if (ok) { done += (size_t)w; left -= (size_t)w; }
    }
  }
  ::close(fd);
  if (!ok) throw std::runtime_error("short write into " + path);
  return 0;
}

int Simulation::timestep_implementation(int /* t */)
{
  B200_CALL(xb_step(ctx, scheme));
  return 0;
}

int Simulation::calculate()
{
  wall_start_ = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  for (int t = start + 1; t <= geom.geom_nt; ++t) {
    if (rank_ == 0) std::cout << std::format("Timestep = {:.4f} [1/w_pe] = {} [dt]", t * geom.dt, t) << "\n";
    for (const Command& cmd : step_presets_)  // interfaces/simulation.cpp:82-84
      if (execute_command(cmd, t)) return 1;
    if (timestep_implementation(t)) return 1;
    if (scheme == XB_ECCAPFIM) {  // the LOG lines of calc_iteration, eccapfim/simulation.cpp:80-96
      int32_t its = 0, fev = 0, reason = 0;
      double fn = 0, cn = 0, tc = 0;
      xb_nonlinear_info(ctx, &its, &fev, &reason, &fn, &cn, &tc);
      std::cout << std::format("  SNESSolve() has finished: reason {}, iterations {}, function evaluations {}, function norm {:e}", reason, its, fev, fn) << "\n";
      std::cout << std::format("    Number of Crank-Nicolson iterations is {:3.4f}", cn) << "\n";
      std::cout << std::format("    Number of traversed cells is {:3.4f}", tc) << "\n";
      if (diagnose_convergence(t)) return 1;
    }
    else {
      int its = 0, reason = 0;
      double rn = 0;
      xb_solver_info(ctx, XB_SOLVER_PREDICT, &its, &rn, &reason);
      std::cout << std::format("  KSPSolve() has finished: reason {}, iterations {}, residual norm {:.3e}", reason, its, rn) << "\n";
    }
    if (charge_ && diagnose_charge(t)) return 1;
    if (diagnose_momentum(t)) return 1;
    if (diagnose_fields(t)) return 1;
    if (diagnose_energy(t)) return 1;
    if (log_levels_ && diagnose_log(t)) return 1;
    if (backup_period_ > 0 && t % backup_period_ == 0 && save_backup(t)) return 1;
  }
  std::cout << "Summary of Stages:\n";  // utils/sync_clock.cpp:85-91
  const char* names_ec[XB_STAGE_COUNT] = {"Clear sources", "First push", "Advance field", "Second push", "Correct fields", "Final update"};
  const char* names_cap[XB_STAGE_COUNT] = {"init_iteration", "-", "calc_iteration", "-", "-", "after_iteration"};  // eccapfim/simulation.cpp:46,72,106
  const char* const* names = scheme == XB_ECCAPFIM ? names_cap : names_ec;
  for (int s = 0; s < XB_STAGE_COUNT; ++s) {
    if (names[s][0] == '-') continue;
    double sec = 0;
    int64_t calls = 0;
    xb_timing(ctx, s, &sec, &calls);
    std::cout << std::format("  {:<16s} {:10.4e} [sec] over {} calls", names[s], sec, calls) << "\n";
  }
  return 0;
}

// LogView (src/diagnostics/log_view.cpp): the per-stage clocks of the step.  The reference reads PETSc's log
// stages; here the stage clocks are CUDA-event times of the library (xb_timing), one "rank" per GPU, so the
// Comm-Avg column is the stage time itself.  Same files, same columns:
//   log-EachTimestep.txt       (:33-110)  Timestep  Total_[sec]  <stage>: seconds and % of the step
//   log-DiagnosePeriodAvg.txt  (:112-235) rewritten every diagnose_period: totals and per-step averages of the period
//   log-AllTimestepsSummary.txt (:237-247) PetscLogView's role: the totals since the start
int Simulation::diagnose_log(int t)
{
  static const char* names_ec[XB_STAGE_COUNT] = {"Clear sources", "First push", "Advance field", "Second push", "Correct fields", "Final update"};
  static const char* names_corr[XB_STAGE_COUNT] = {"Clear sources", "First push", "Predict field", "Second push", "Correct fields", "Final update"};
  static const char* names_cap[XB_STAGE_COUNT] = {"Init iteration", "-", "Calc iteration", "-", "-", "After iteration"};
  const char* const* names = scheme == XB_ECCAPFIM ? names_cap : (scheme == XB_ECSIMCORR ? names_corr : names_ec);
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  if (log_prev_wall_ == 0) log_prev_wall_ = log_period_wall_ = wall_start_;
  std::array<double, XB_STAGE_COUNT> now{};
  for (int st = 0; st < XB_STAGE_COUNT; ++st) {
    int64_t calls = 0;
    B200_CALL(xb_timing(ctx, st, &now[st], &calls));
  }
  if (rank_ != 0) {
    log_prev_wall_ = wall;
    log_prev_stage_ = now;
    return 0;
  }
  auto underscored = [](std::string v) {
    for (char& ch : v)
      if (ch == ' ') ch = '_';
    return v;
  };
  if (log_levels_ & 1) {
    if (!log_each_) {
      log_each_ = std::make_unique<std::ofstream>(out_dir + "/log-EachTimestep.txt");
      *log_each_ << "Timestep  Total_[sec]  " << std::format("{:<19s}", "Main_Stage");
      for (int st = 0; st < XB_STAGE_COUNT; ++st)
        if (names[st][0] != '-') *log_each_ << std::format("{:<19s}", underscored(names[st]));
      *log_each_ << "\n";
    }
    const double total = wall - log_prev_wall_;
    double staged = 0.0;
    for (int st = 0; st < XB_STAGE_COUNT; ++st) staged += now[st] - log_prev_stage_[st];
    *log_each_ << std::format("{:5d}     {:9.3e}    ", t, total);
    auto cell = [&](double sec) { *log_each_ << std::format("{:<6.4e} {:5.1f}%  ", sec, total > 0.0 ? 100.0 * sec / total : 0.0); };
    cell(std::max(0.0, total - staged));  // PETSc's "Main Stage": everything outside the logged stages (diagnostics, I/O)
    for (int st = 0; st < XB_STAGE_COUNT; ++st)
      if (names[st][0] != '-') cell(now[st] - log_prev_stage_[st]);
    *log_each_ << "\n";
    if (t % geom.diagnose_period == 0) log_each_->flush();
  }
  if (t % geom.diagnose_period == 0) {
    auto summary = [&](const std::string& file, double t0, const std::array<double, XB_STAGE_COUNT>& s0, int steps) {
      std::ofstream f(file);
      const double period = wall - t0;
      f << std::format("Total time (sec):     {:5.3e}\n", period);
      f << "\nSummary of Stages:    ------------ Time ------------\n";
      f << "                        Comm-Avg  Period-Avg  %Total\n";
      double staged = 0.0;
      for (int st = 0; st < XB_STAGE_COUNT; ++st) staged += now[st] - s0[st];
      auto line = [&](int idx, const char* name, double sec) {
        if (period <= 0.0 || 100.0 * sec / period < 0.1) return;
        f << std::format("{:2d}:  {:>15s}: {:6.4e}  {:6.4e}  {:5.1f}%\n", idx, name, sec, sec / std::max(steps, 1), 100.0 * sec / period);
      };
      line(0, "Main Stage", std::max(0.0, period - staged));
      for (int st = 0, idx = 1; st < XB_STAGE_COUNT; ++st)
        if (names[st][0] != '-') line(idx++, names[st], now[st] - s0[st]);
      f << "\n";
    };
    if (log_levels_ & 2) summary(out_dir + "/log-DiagnosePeriodAvg.txt", log_period_wall_, log_period_stage_, geom.diagnose_period);
    if (log_levels_ & 4) summary(out_dir + "/log-AllTimestepsSummary.txt", wall_start_, std::array<double, XB_STAGE_COUNT>{}, t - start);
    log_period_wall_ = wall;
    log_period_stage_ = now;
  }
  log_prev_wall_ = wall;
  log_prev_stage_ = now;
  return 0;
}

int Simulation::finalize()
{
  if (log_each_) log_each_->flush();
  if (energy_) energy_->flush();
  if (energy_cons_) energy_cons_->flush();
  if (convergence_) convergence_->flush();
  if (charge_) charge_->flush();
  if (momentum_) momentum_->flush();
  if (ctx) {
    xb_destroy(ctx);
    ctx = nullptr;
  }
  return 0;
}

// ---- SimulationBackup (src/diagnostics/simulation_backup.cpp:27-183) ---------------------------------------
// Fields: PETSc's binary Vec image (VecView through a binary viewer): int32 VEC_FILE_CLASSID = 1211214,
// int32 length, then the entries as float64 -- all big-endian, natural [z][y][x][c] order for a DMDA vector --
// plus the `.info` side file PETSc writes for a blocked vector.  Particles: `<sort>.numparts` = one int32 and
// `<sort>` = 6 float64 per particle in storage order, header-less (PetscViewerBinarySetSkipHeader), big-endian.
// PETSc is not available here and the reference ships no backup file: the layout follows PETSc's documented
// binary format and is checked by a save / restart round trip only (tests/test_gpu_parity.py).
namespace {

template <class T>
T byteswap(T v)
{
  unsigned char* p = reinterpret_cast<unsigned char*>(&v);
  for (size_t i = 0; i < sizeof(T) / 2; ++i) std::swap(p[i], p[sizeof(T) - 1 - i]);
  return v;
}

void write_be_doubles(std::ofstream& f, const double* v, size_t n)
{
  std::vector<double> be(v, v + n);
  for (double& x : be) x = byteswap(x);
  f.write(reinterpret_cast<const char*>(be.data()), (std::streamsize)(n * sizeof(double)));
}

void read_be_doubles(std::ifstream& f, double* v, size_t n)
{
  f.read(reinterpret_cast<char*>(v), (std::streamsize)(n * sizeof(double)));
  for (size_t i = 0; i < n; ++i) v[i] = byteswap(v[i]);
}

constexpr int32_t VEC_FILE_CLASSID = 1211214;

}  // namespace

int Simulation::save_backup(int t)
{
  const std::string dir = std::format("{}/simulation_backup/{}", out_dir, t);
  std::filesystem::create_directories(dir);
  std::vector<double> f;
  const size_t n3 = (size_t)3 * geom.geom_nx * geom.geom_ny * geom.geom_nz;
  for (const char* name : {"E", "B", "B0"}) {  // simulation_backup_builder.cpp:17-21
    if (get_named_vector(name, f)) return 1;
    const std::string path = dir + "/" + name;
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT, 0644);
    if (fd < 0) throw std::runtime_error("SimulationBackup: cannot write " + path);
    bool ok = true;
    if (rank_ == 0) {
      const int32_t header[2] = {byteswap(VEC_FILE_CLASSID), byteswap((int32_t)n3)};
      ok = ::pwrite(fd, header, sizeof(header), 0) == (ssize_t)sizeof(header);
      std::ofstream(path + ".info") << "-vecload_block_size 3\n";
    }
    // the slab of this rank inside the natural-order image (one file for all ranks, as VecView writes it)
    std::vector<double> be(f);
    for (double& x : be) x = byteswap(x);
    const off_t at = 8 + (off_t)sizeof(double) * 3 * geom.geom_nx * geom.geom_ny * (off_t)z0_;
    const char* data = reinterpret_cast<const char*>(be.data());
    size_t left = be.size() * sizeof(double), done = 0;
    while (ok && left > 0) {
      const ssize_t w = ::pwrite(fd, data + done, left, at + (off_t)done);
      ok = w > 0;
      if (ok) { done += (size_t)w; left -= (size_t)w; }
    }
    ::close(fd);
    if (!ok) throw std::runtime_error("SimulationBackup: cannot write " + path);
  }
  std::vector<Point> pts;
  for (auto& sort : particles_) {
    if (sort->download(pts)) return 1;
    // one rank: the reference's file names; several ranks: one pair of files per rank (the reference's MPI-IO viewer
    // concatenates the ranks' particles into one file, which needs the counts of the lower ranks first)
    const std::string base = dir + "/" + sort->parameters.sort_name + (nranks_ > 1 ? std::format(".rank{}", rank_) : std::string());
    const int32_t n = byteswap((int32_t)pts.size());
    std::ofstream(base + ".numparts", std::ios::binary).write(reinterpret_cast<const char*>(&n), sizeof(n));
    std::ofstream file(base, std::ios::binary);
    if (!pts.empty()) write_be_doubles(file, &pts[0].r[0], 6 * pts.size());
    if (!file) throw std::runtime_error("SimulationBackup: cannot write " + base);
  }
  // save_temporal_diagnostics (:98-106): the tables as they are now
  for (Table* tb : {energy_.get(), energy_cons_.get(), convergence_.get(), charge_.get(), momentum_.get()})
    if (tb) tb->flush();
  if (rank_ != 0) return 0;
  if (std::filesystem::exists(out_dir + "/temporal"))
    std::filesystem::copy(out_dir + "/temporal", dir + "/temporal",
                          std::filesystem::copy_options::overwrite_existing | std::filesystem::copy_options::recursive);
  // only the last two periods are kept (num_periods_being_kept, simulation_backup.h:46)
  std::filesystem::remove_all(std::format("{}/simulation_backup/{}", out_dir, t - 2 * backup_period_));
  return 0;
}

int Simulation::load_backup(int t)
{
  const std::string dir = std::format("{}/simulation_backup/{}", out_dir, t);
  if (!std::filesystem::exists(dir)) throw std::runtime_error("Cannot load the timestep, no backup directory " + dir);
  const size_t n3 = (size_t)3 * geom.geom_nx * geom.geom_ny * geom.geom_nz;
  const size_t nloc = (size_t)3 * geom.geom_nx * geom.geom_ny * nzl_, off = (size_t)3 * geom.geom_nx * geom.geom_ny * z0_;
  std::vector<double> f(nloc);
  const std::pair<const char*, int> fields[] = {{"E", XB_E}, {"B", XB_B}, {"B0", XB_B0}};
  for (auto& [name, id] : fields) {
    std::ifstream file(dir + "/" + name, std::ios::binary);
    int32_t header[2] = {0, 0};
    file.read(reinterpret_cast<char*>(header), sizeof(header));
    if (!file || byteswap(header[0]) != VEC_FILE_CLASSID || (size_t)byteswap(header[1]) != n3)
      throw std::runtime_error(std::string("SimulationBackup: ") + name + " is not a PETSc binary Vec of this geometry");
    file.seekg((std::streamoff)(8 + off * sizeof(double)));
    read_be_doubles(file, f.data(), nloc);
    if (!file) throw std::runtime_error(std::string("SimulationBackup: short read of ") + name);
    B200_CALL(xb_field_upload(ctx, id, 0, f.data()));
  }
  for (auto& sort : particles_) {
    // every rank reads every particle file and keeps the particles of its slab (add_particle)
    std::vector<std::string> bases;
    const std::string plain = dir + "/" + sort->parameters.sort_name;
    if (std::filesystem::exists(plain + ".numparts")) bases.push_back(plain);
    for (int r = 0; std::filesystem::exists(std::format("{}.rank{}.numparts", plain, r)); ++r) bases.push_back(std::format("{}.rank{}", plain, r));
    if (bases.empty()) throw std::runtime_error("SimulationBackup: no particle file " + plain);
    for (const std::string& base : bases) {
      int32_t n = 0;
      std::ifstream(base + ".numparts", std::ios::binary).read(reinterpret_cast<char*>(&n), sizeof(n));
      n = byteswap(n);
      std::ifstream file(base, std::ios::binary);
      std::vector<double> raw((size_t)6 * (size_t)std::max(n, 0));
      read_be_doubles(file, raw.data(), raw.size());
      if (!file) throw std::runtime_error("SimulationBackup: short read of " + base);
      for (int32_t i = 0; i < n; ++i) {
        Point pt;
        for (int c = 0; c < 3; ++c) {
          pt.r[c] = raw[6 * (size_t)i + c];
          pt.p[c] = raw[6 * (size_t)i + 3 + c];
        }
        sort->add_particle(pt);
      }
    }
    if (sort->flush()) return 1;
  }
  if (rank_ == 0 && std::filesystem::exists(dir + "/temporal"))  // load_temporal_diagnostics (:162-169)
    std::filesystem::copy(dir + "/temporal", out_dir + "/temporal",
                          std::filesystem::copy_options::overwrite_existing | std::filesystem::copy_options::recursive);
  std::cout << std::format("  Simulation is successfully loaded from {:.1f} [1/w_pe], {} [dt]", t * geom.dt, t) << "\n";
  return 0;
}

// MomentumConservation::add_columns (src/diagnostics/momentum_conservation.cpp:29-68), sums on the device
int Simulation::diagnose_momentum(int t)
{
  auto num = [](double v) { return std::format("{: .6e}", v); };
  auto len = [](const std::array<double, 3>& a) { return std::hypot(a[0], a[1], a[2]); };
  momentum_->add(6, "Time", std::format("{:d}", t));
  std::array<double, 3> sum = {0.0, 0.0, 0.0};
  for (size_t i = 0; i < particles_.size(); ++i) {
    double o[6];
    B200_CALL(xb_momentum(ctx, (int32_t)i, o));
    const std::string& name = particles_[i]->parameters.sort_name;
    const std::array<double, 3> p1 = {o[0], o[1], o[2]}, qe = {o[3], o[4], o[5]}, p0 = P0_[i];
    for (int c = 0; c < 3; ++c) momentum_->add(13, std::string("P") + "xyz"[c] + "_" + name, num(p1[c]));
    for (int c = 0; c < 3; ++c) momentum_->add(13, std::string("QE") + "xyz"[c] + "_" + name, num(qe[c]));
    std::array<double, 3> err, dp, sp;
    for (int c = 0; c < 3; ++c) {
      err[c] = (p1[c] - p0[c]) / geom.dt - qe[c];
      sum[c] += err[c];
      dp[c] = p1[c] - p0[c];
      sp[c] = p1[c] + p0[c];
    }
    double freq = 0.0;
    if (const double denom = len(sp); std::abs(denom) > 1e-10) freq = (len(dp) / denom) / (0.5 * geom.dt);  // PETSC_SMALL
    momentum_->add(13, "N2dP_" + name, num(len(err)));
    momentum_->add(13, "fP_" + name, num(freq));
    P0_[i] = p1;
  }
  momentum_->add(13, "N2dP", num(len(sum)));
  momentum_->row(t == 0);
  if (t % geom.diagnose_period == 0) momentum_->flush();
  return 0;
}

// ChargeConservation::add_columns (src/diagnostics/charge_conservation.cpp:125-171), evaluated on the device
int Simulation::diagnose_charge(int t)
{
  std::vector<double> norms(2 * (particles_.size() + 1), 0.0);
  B200_CALL(xb_charge_conservation(ctx, scheme == XB_ECSIMCORR ? 0 : 1, norms.data()));
  auto num = [](double v) { return std::format("{: .6e}", v); };
  charge_->add(6, "Time", std::format("{:d}", t));
  for (size_t i = 0; i < particles_.size(); ++i) {
    charge_->add(13, "N1dQ_" + particles_[i]->parameters.sort_name, num(norms[2 * i]));
    charge_->add(13, "N2dQ_" + particles_[i]->parameters.sort_name, num(norms[2 * i + 1]));
  }
  charge_->add(13, "N1dQ_tot", num(norms[2 * particles_.size()]));
  charge_->add(13, "N2dQ_tot", num(norms[2 * particles_.size() + 1]));
  charge_->row(!charge_header_);
  charge_header_ = true;
  if (t % geom.diagnose_period == 0) charge_->flush();
  return 0;
}

// eccapfim::ConvergenceHistory::add_columns (src/impls/eccapfim/convergence_history.cpp:11-44)
int Simulation::diagnose_convergence(int t)
{
  int32_t its = 0, fev = 0, reason = 0, len = 0;
  double fn = 0, cn = 0, tc = 0;
  std::vector<double> hist(2048);
  if (t > 0) {
    B200_CALL(xb_nonlinear_info(ctx, &its, &fev, &reason, &fn, &cn, &tc));
    B200_CALL(xb_nonlinear_history(ctx, hist.data(), (int32_t)hist.size(), &len));
  }
  convergence_->add(6, "Time", std::format("{:d}", t));
  for (auto& p : particles_) {  // the averages are over all sorts here; the reference prints them per sort
    convergence_->add(8, "AvgCN_" + p->parameters.sort_name, std::format("{:.3f}", cn));
    convergence_->add(8, "AvgTC_" + p->parameters.sort_name, std::format("{:.3f}", tc));
  }
  convergence_->add(6, "FEvals", std::format("{:d}", fev));
  convergence_->add(6, "ItNum", std::format("{:d}", its));
  if (len == 0) convergence_->add(12, "ConvHist", "");
  for (int i = 0; i < len && i < (int)hist.size(); ++i) convergence_->add(12, "ConvHist", std::format("{:8.6e}", hist[i]));
  convergence_->row(t == 0);
  if (t % geom.diagnose_period == 0) convergence_->flush();
  return 0;
}

// after a restart: the energies of t = start become the "previous" values, nothing is written
int Simulation::prime_energy()
{
  energy_silent_ = true;
  const int rc = diagnose_energy(start);
  energy_silent_ = false;
  return rc;
}

// Energy::diagnose (src/diagnostics/energy.cpp:20-180) + ecsimcorr::Energy (ecsimcorr/simulation.cpp:169-197)
int Simulation::diagnose_energy(int t)
{
  auto field = [&](const char* name, double& w, double& sd) {
    // VecNorm and VecStrideSumAll of energy.cpp:43-59 on the device, summed over all ranks
    double sums[4];
    B200_CALL(xb_field_sums(ctx, std::string(name) == "E" ? XB_E : XB_B, 0, sums));
    w = 0.5 * sums[3];  // energy.cpp:46-50 (the norm is squared again there)
    const double g3 = (double)geom.geom_nx * geom.geom_ny * geom.geom_nz;
    sd = std::sqrt((w - 0.5 * (sums[0] * sums[0] + sums[1] * sums[1] + sums[2] * sums[2]) / g3) / g3);
    return 0;
  };
  auto kinetic = [&]() {
    for (size_t i = 0; i < particles_.size(); ++i) {
      double mo[5];
      B200_CALL(xb_particle_moments(ctx, (int32_t)i, mo));
      const SortParameters& sp = particles_[i]->parameters;
      const double fr = 0.5 * sp.m * (sp.n / (double)sp.Np);
      K_[i] = fr * mo[3];
      const double s = mo[3] - (mo[0] * mo[0] + mo[1] * mo[1] + mo[2] * mo[2]) / mo[4];
      stdK_[i] = mo[4] > 0 ? fr * std::sqrt(std::abs(s) / mo[4]) : 0.0;
    }
    return 0;
  };
  if (t == 0) {
    if (field("E", E_, stdE_) || field("B", B_, stdB_) || kinetic()) return 1;
  }
  E0_ = E_;
  B0_ = B_;
  K0_ = K_;
  if (field("E", E_, stdE_) || field("B", B_, stdB_) || kinetic()) return 1;
  if (energy_silent_) return 0;

  auto num = [](double v) { return std::format("{: .6e}", v); };
  energy_->add(6, "Time", std::format("{:d}", t));
  energy_->add(13, "wE", num(E_));
  energy_->add(13, "wB", num(B_));
  for (size_t i = 0; i < particles_.size(); ++i) energy_->add(13, "wK_" + particles_[i]->parameters.sort_name, num(K_[i]));
  energy_->add(13, "sE", num(stdE_));
  energy_->add(13, "sB", num(stdB_));
  for (size_t i = 0; i < particles_.size(); ++i) energy_->add(13, "sK_" + particles_[i]->parameters.sort_name, num(stdK_[i]));
  energy_->row(t == 0);

  const double dE = E_ - E0_, dB = B_ - B0_;
  double dK = 0.0;
  energy_cons_->add(6, "Time", std::format("{:d}", t));
  energy_cons_->add(13, "dE", num(dE));
  energy_cons_->add(13, "dB", num(dB));
  for (size_t i = 0; i < particles_.size(); ++i) {
    const std::string& name = particles_[i]->parameters.sort_name;
    energy_cons_->add(13, "dK_" + name, num(K_[i] - K0_[i]));
    dK += K_[i] - K0_[i];
    if (scheme == XB_ECSIMCORR) {
      Particles& p = *particles_[i];
      energy_cons_->add(13, "CWD_" + name, num(p.scalar(XB_LAMBDA_DK)));
      energy_cons_->add(13, "PWD_" + name, num(p.scalar(XB_PRED_DK) - geom.dt * p.scalar(XB_PRED_W)));
      energy_cons_->add(13, "LdK_" + name, num(p.scalar(XB_CORR_DK) - geom.dt * p.scalar(XB_CORR_W)));
    }
  }
  energy_cons_->add(13, "dE+dB+dK", num(dE + dB + dK));
  if (scheme == XB_ECSIMCORR) {
    double corr_w = 0.0;
    for (auto& p : particles_) corr_w += p->scalar(XB_CORR_W);
    energy_cons_->add(13, "WD", num(dK - geom.dt * corr_w));
  }
  energy_cons_->row(t == 0);
  if (t % geom.diagnose_period == 0) {
    energy_->flush();
    energy_cons_->flush();
  }
  return 0;
}

}  // namespace b200
