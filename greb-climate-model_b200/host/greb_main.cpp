// greb — the compiled host of the B200 path: what PROGRAM greb_run does (reference src/greb.f90:996-1098),
// written against the C ABI of include/greb_b200.h only (no CUDA, no torch, no Python in this file).
//
//   greb [namelist ...]            one member per namelist, all of them stepped as ONE ensemble on the GPU
//   greb --check [namelist ...]    parse only: print what would run (one JSON line per namelist), no GPU
//   greb --original [namelist_original]   the greb-original driver (control run + scenario, `log_exp` experiments)
//   options: --input DIR (default "input"), --device N, --gpus G (members in contiguous blocks over devices
//            N..N+G-1, one handle and one host thread per GPU), --arith exact|fast
//
// Like the reference: with no argument the file `namelist` is read (f:1030-1037); the four groups
// physics_par, numerics_par, diagnostics_par, co2_par set the member's parameters (f:1040-1048); co2_ppm is
// padded by the rule of f:1050-1061; the ten direct-access files of f:1018-1027/1073-1085 are read from the
// input directory; every simulated year prints the console line of f:954 and the monthly means go to
// output_file[_ens_id] as 5 records per month (f:978-982).  Unlike the reference, several namelists may be
// given: the members share the inputs and must share time_flux and time_scnr (one launch advances all).
// The Python twin of this file (the one the tests drive) is greb_b200/host.py; both produce the same bytes.
//
// Exit codes: 0 ok, 2 usage / namelist / input error, 3 the library reported an error (message on stderr).
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

#include "greb_b200.h"

namespace {

struct NamelistError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- values ------------------------------------------------------------------------------------
struct Value {
  enum Kind { NONE, STRING, LOGICAL, INTEGER, REAL } kind = NONE;   // NONE = position never assigned
  std::string s;
  long i = 0;
  double r = 0;
  double number(const std::string& name) const {
    if (kind == INTEGER) return (double)i;
    if (kind == REAL) return r;
    throw NamelistError(name + ": a number is expected");
  }
};

std::string lower(std::string t) {
  for (auto& c : t) c = (char)std::tolower((unsigned char)c);
  return t;
}
std::string trim(const std::string& t) {
  size_t a = 0, b = t.size();
  while (a < b && std::isspace((unsigned char)t[a])) ++a;
  while (b > a && std::isspace((unsigned char)t[b - 1])) --b;
  return t.substr(a, b - a);
}

Value scalar(const std::string& tok) {
  const std::string t = trim(tok);
  if (t.empty()) throw NamelistError("empty value");
  Value v;
  if (t[0] == '\'' || t[0] == '"') {
    if (t.size() < 2 || t.back() != t[0]) throw NamelistError("unterminated string " + t);
    v.kind = Value::STRING;
    v.s = t.substr(1, t.size() - 2);
    return v;
  }
  const std::string tl = lower(t);
  if (tl == ".true." || tl == "t" || tl == ".t.") {
    v.kind = Value::LOGICAL;
    v.i = 1;
    return v;
  }
  if (tl == ".false." || tl == "f" || tl == ".f.") {
    v.kind = Value::LOGICAL;
    return v;
  }
  size_t k = (t[0] == '+' || t[0] == '-') ? 1 : 0;
  bool digits = k < t.size();
  for (size_t j = k; j < t.size(); ++j) digits = digits && std::isdigit((unsigned char)t[j]);
  if (digits) {
    v.kind = Value::INTEGER;
    v.i = std::strtol(t.c_str(), nullptr, 10);
    return v;
  }
  std::string num = tl;
  for (auto& c : num)
    if (c == 'd') c = 'e';                       // Fortran double-precision exponent
  char* end = nullptr;
  errno = 0;
  const double r = std::strtod(num.c_str(), &end);
  if (end == num.c_str() || *end != 0) throw NamelistError("cannot parse value '" + t + "'");
  v.kind = Value::REAL;
  v.r = r;
  return v;
}

// "1, 2 3, 4*0.5, 'a b'" or "(/ 1, 2 /)": separators are commas and blanks outside quotes, r*c repeats c
std::vector<Value> values(std::string text) {
  text = trim(text);
  while (!text.empty() && text.back() == ',') text = trim(text.substr(0, text.size() - 1));
  if (text.size() >= 4 && text.compare(0, 2, "(/") == 0 && text.compare(text.size() - 2, 2, "/)") == 0)
    text = text.substr(2, text.size() - 4);
  std::vector<std::string> parts;
  std::string cur;
  char q = 0;
  for (char ch : text) {
    if (q) {
      cur.push_back(ch);
      if (ch == q) q = 0;
    } else if (ch == '\'' || ch == '"') {
      q = ch;
      cur.push_back(ch);
    } else if (ch == ',' || (std::isspace((unsigned char)ch) && !trim(cur).empty())) {
      if (!trim(cur).empty()) parts.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(ch);
    }
  }
  if (!trim(cur).empty()) parts.push_back(cur);
  std::vector<Value> out;
  for (const auto& p0 : parts) {
    const std::string p = trim(p0);
    size_t star = std::string::npos;
    if (!p.empty() && std::isdigit((unsigned char)p[0])) {
      size_t j = 0;
      while (j < p.size() && std::isdigit((unsigned char)p[j])) ++j;
      if (j < p.size() && p[j] == '*' && j + 1 < p.size()) star = j;
    }
    if (star != std::string::npos) {
      const long n = std::strtol(p.substr(0, star).c_str(), nullptr, 10);
      const Value v = scalar(p.substr(star + 1));
      for (long k = 0; k < n; ++k) out.push_back(v);
    } else {
      out.push_back(scalar(p));
    }
  }
  return out;
}

std::string strip_comment(const std::string& line) {
  std::string out;
  char q = 0;
  for (char ch : line) {
    if (q) {
      out.push_back(ch);
      if (ch == q) q = 0;
    } else if (ch == '\'' || ch == '"') {
      q = ch;
      out.push_back(ch);
    } else if (ch == '!') {
      break;
    } else {
      out.push_back(ch);
    }
  }
  return out;
}

bool ident_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
bool ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

using Group = std::map<std::string, std::vector<Value>>;

// {group: {name: values}} — names lower-cased; `a(3) = 5, 6` fills elements 3 and 4, the others stay NONE
std::map<std::string, Group> parse_namelist(const std::string& text) {
  std::string body;
  {
    std::istringstream in(text);
    std::string ln;
    bool first = true;
    while (std::getline(in, ln)) {
      if (!first) body.push_back('\n');
      first = false;
      body += strip_comment(ln);
    }
  }
  std::map<std::string, Group> groups;
  size_t pos = 0;
  while (true) {
    size_t amp = std::string::npos;
    size_t name_b = 0, name_e = 0;
    for (size_t i = pos; i < body.size(); ++i) {
      if (body[i] != '&') continue;
      size_t j = i + 1;
      while (j < body.size() && std::isspace((unsigned char)body[j])) ++j;
      if (j < body.size() && ident_start(body[j])) {
        amp = i;
        name_b = j;
        name_e = j;
        while (name_e < body.size() && ident_char(body[name_e])) ++name_e;
        break;
      }
    }
    if (amp == std::string::npos) break;
    const std::string gname = lower(body.substr(name_b, name_e - name_b));
    const size_t start = name_e;
    size_t end = std::string::npos;
    char q = 0;
    for (size_t i = start; i < body.size(); ++i) {
      const char ch = body[i];
      if (q) {
        if (ch == q) q = 0;
      } else if (ch == '\'' || ch == '"') {
        q = ch;
      } else if (ch == '/' && !((i > 0 && body[i - 1] == '(') || (i + 1 < body.size() && body[i + 1] == ')'))) {
        end = i;
        break;
      } else if (lower(body.substr(i, 4)) == "&end") {
        end = i;
        break;
      }
    }
    if (end == std::string::npos) throw NamelistError("namelist group &" + gname + " is not terminated by '/'");
    const std::string content = body.substr(start, end - start);
    std::string masked = content;   // quoted text may hold '=' or '(1)': blank it for the pair search only
    q = 0;
    for (auto& ch : masked) {
      if (q) {
        if (ch == q) q = 0;
        else ch = '_';
      } else if (ch == '\'' || ch == '"') {
        q = ch;
      }
    }
    // pairs "name =" / "name(i) =": an identifier (not inside a longer token), optional subscript, '='
    struct Pair {
      std::string name;
      long first;
      size_t begin, value_begin;
    };
    std::vector<Pair> pairs;
    for (size_t i = 0; i < masked.size();) {
      if (!ident_start(masked[i]) || (i > 0 && ident_char(masked[i - 1]))) {
        ++i;
        continue;
      }
      size_t j = i;
      while (j < masked.size() && ident_char(masked[j])) ++j;
      size_t k = j;
      while (k < masked.size() && std::isspace((unsigned char)masked[k])) ++k;
      long first = 0;
      bool has_sub = false;
      if (k < masked.size() && masked[k] == '(') {
        size_t m = k + 1;
        while (m < masked.size() && std::isspace((unsigned char)masked[m])) ++m;
        size_t d = m;
        while (d < masked.size() && std::isdigit((unsigned char)masked[d])) ++d;
        size_t c = d;
        while (c < masked.size() && std::isspace((unsigned char)masked[c])) ++c;
        if (d > m && c < masked.size() && masked[c] == ')') {
          first = std::strtol(masked.substr(m, d - m).c_str(), nullptr, 10) - 1;
          has_sub = true;
          k = c + 1;
          while (k < masked.size() && std::isspace((unsigned char)masked[k])) ++k;
        }
      }
      if (k < masked.size() && masked[k] == '=') {
        if (has_sub && first < 0) throw NamelistError(masked.substr(i, j - i) + ": subscripts start at 1");
        pairs.push_back({lower(masked.substr(i, j - i)), first, i, k + 1});
        i = k + 1;
      } else {
        i = j;
      }
    }
    Group items;
    std::map<std::string, std::map<long, Value>> slots;
    for (size_t p = 0; p < pairs.size(); ++p) {
      const size_t v_end = p + 1 < pairs.size() ? pairs[p + 1].begin : content.size();
      const auto vals = values(content.substr(pairs[p].value_begin, v_end - pairs[p].value_begin));
      auto& d = slots[pairs[p].name];
      for (size_t k = 0; k < vals.size(); ++k) d[pairs[p].first + (long)k] = vals[k];
    }
    for (auto& kv : slots) {
      std::vector<Value> v;
      if (!kv.second.empty()) {
        v.resize((size_t)kv.second.rbegin()->first + 1);
        for (auto& e : kv.second) v[(size_t)e.first] = e.second;
      }
      items[kv.first] = v;
    }
    if (groups.count(gname)) throw NamelistError("namelist group &" + gname + " appears twice");
    groups[gname] = items;
    pos = end + 1;
  }
  return groups;
}

// ---- one ./greb invocation (f:1042-1068) ---------------------------------------------------------
struct RunConfig {
  greb_physics_par p;
  int time_flux = 0, time_scnr = 0, year0 = 1940, ipx = 1, ipy = 1;
  std::string output_file = "output/scenario", ens_id;
  std::vector<float> co2_ppm;
  std::string output_file_full() const {                              // f:1063-1068
    const std::string e = trim(ens_id);
    return e.empty() ? output_file : trim(output_file) + "_" + e;
  }
};

const Value& only(const std::string& name, const std::vector<Value>& v) {
  if (v.size() != 1 || v[0].kind == Value::NONE)
    throw NamelistError(name + " is a scalar: exactly one value, no subscript");
  return v[0];
}

struct PhysField {
  const char* name;
  float greb_physics_par::*ptr;
};
const PhysField kPhys[] = {
    {"pi", &greb_physics_par::pi}, {"sig", &greb_physics_par::sig}, {"rho_ocean", &greb_physics_par::rho_ocean},
    {"rho_land", &greb_physics_par::rho_land}, {"rho_air", &greb_physics_par::rho_air},
    {"cp_ocean", &greb_physics_par::cp_ocean}, {"cp_land", &greb_physics_par::cp_land},
    {"cp_air", &greb_physics_par::cp_air}, {"eps", &greb_physics_par::eps}, {"d_ocean", &greb_physics_par::d_ocean},
    {"d_land", &greb_physics_par::d_land}, {"d_air", &greb_physics_par::d_air}, {"ct_sens", &greb_physics_par::ct_sens},
    {"da_ice", &greb_physics_par::da_ice}, {"a_no_ice", &greb_physics_par::a_no_ice},
    {"a_cloud", &greb_physics_par::a_cloud}, {"tl_ice1", &greb_physics_par::Tl_ice1},
    {"tl_ice2", &greb_physics_par::Tl_ice2}, {"to_ice1", &greb_physics_par::To_ice1},
    {"to_ice2", &greb_physics_par::To_ice2}, {"co_turb", &greb_physics_par::co_turb},
    {"kappa", &greb_physics_par::kappa}, {"ce", &greb_physics_par::ce}, {"cq_latent", &greb_physics_par::cq_latent},
    {"cq_rain", &greb_physics_par::cq_rain}, {"z_air", &greb_physics_par::z_air},
    {"z_vapor", &greb_physics_par::z_vapor}, {"r_qviwv", &greb_physics_par::r_qviwv},
};

RunConfig config_from_namelist(const std::string& text) {
  auto g = parse_namelist(text);
  for (auto& kv : g)
    if (kv.first != "physics_par" && kv.first != "numerics_par" && kv.first != "diagnostics_par" && kv.first != "co2_par")
      throw NamelistError("unknown namelist group: " + kv.first);
  RunConfig c;
  greb_b200_physics_defaults(&c.p);
  for (auto& kv : g["physics_par"]) {
    const std::string& k = kv.first;
    if (k == "p_emi") {
      if (kv.second.size() > 10) throw NamelistError("p_emi has 10 elements");
      for (size_t i = 0; i < kv.second.size(); ++i)
        if (kv.second[i].kind != Value::NONE) c.p.p_emi[i] = (float)kv.second[i].number(k);
      continue;
    }
    bool found = false;
    for (const auto& f : kPhys)
      if (k == f.name) {
        c.p.*(f.ptr) = (float)only(k, kv.second).number(k);
        found = true;
      }
    if (!found) throw NamelistError("physics_par: unknown variable " + k);
  }
  for (auto& kv : g["numerics_par"]) {
    const std::string& k = kv.first;
    int* dst = k == "ipx" ? &c.ipx : k == "ipy" ? &c.ipy : k == "time_flux" ? &c.time_flux
             : k == "time_scnr" ? &c.time_scnr : k == "year0" ? &c.year0 : nullptr;
    if (!dst) throw NamelistError("numerics_par: unknown variable " + k);
    *dst = (int)only(k, kv.second).number(k);
  }
  for (auto& kv : g["diagnostics_par"]) {
    const std::string& k = kv.first;
    if (k != "output_file" && k != "ens_id") throw NamelistError("diagnostics_par: unknown variable " + k);
    const Value& v = only(k, kv.second);
    std::string s = v.kind == Value::STRING ? v.s : v.kind == Value::INTEGER ? std::to_string(v.i) : std::string();
    if (v.kind != Value::STRING && v.kind != Value::INTEGER) throw NamelistError(k + ": a string is expected");
    (k == "output_file" ? c.output_file : c.ens_id) = s;
  }
  std::vector<float> given;
  for (auto& kv : g["co2_par"]) {
    const std::string& k = kv.first;
    if (k == "co2_flux") c.p.co2_flux = (float)only(k, kv.second).number(k);
    else if (k == "co2_ppm")
      for (const auto& v : kv.second) given.push_back(v.kind == Value::NONE ? -1.0f : (float)v.number(k));   // f:1047
    else throw NamelistError("co2_par: unknown variable " + k);
  }
  const int ny = c.time_scnr > 0 ? c.time_scnr : 0;
  if ((int)given.size() > ny)   // gfortran aborts when more values than allocated elements are supplied
    throw NamelistError("co2_ppm has " + std::to_string(given.size()) + " values but time_scnr = " + std::to_string(c.time_scnr));
  c.co2_ppm.assign((size_t)ny, -1.0f);
  if (ny > 0) greb_b200_pad_co2(given.data(), (int)given.size(), c.co2_ppm.data(), ny);
  return c;
}

std::string read_text(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw NamelistError("cannot open namelist file '" + path + "'");
  std::ostringstream ss;
  ss << in.rdbuf();
  return ss.str();
}

// ---- files ---------------------------------------------------------------------------------------
std::vector<float> read_field(const std::string& dir, const char* name, size_t n_floats) {
  const std::string path = dir + "/" + name;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error(path + ": cannot open input file");
  std::vector<float> a(n_floats);
  const size_t got = std::fread(a.data(), sizeof(float), n_floats, f);
  std::fclose(f);
  if (got != n_floats)
    throw std::runtime_error(path + ": " + std::to_string(got * 4) + " bytes, expected " + std::to_string(n_floats * 4));
  return a;
}

void make_dirs(const std::string& path) {   // mkdir -p of the directory part
  for (size_t i = 1; i < path.size(); ++i)
    if (path[i] == '/') ::mkdir(path.substr(0, i).c_str(), 0777);
}

// the reference opens the output without status='replace' (f:174): a longer existing file is not truncated
void write_output(const std::string& path, const float* monthly, size_t n_floats) {
  make_dirs(path);
  FILE* f = std::fopen(path.c_str(), "r+b");
  if (!f) f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error(path + ": cannot open output file");
  std::fseek(f, 0, SEEK_SET);
  const size_t put = std::fwrite(monthly, sizeof(float), n_floats, f);
  std::fclose(f);
  if (put != n_floats) throw std::runtime_error(path + ": short write");
}

std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') o.push_back('\\');
    o.push_back(c);
  }
  return o;
}

void print_check(const std::string& file, const RunConfig& c) {
  std::printf("{\"namelist\": \"%s\", \"time_flux\": %d, \"time_scnr\": %d, \"year0\": %d, \"ipx\": %d, \"ipy\": %d, "
              "\"output_file_full\": \"%s\", \"physics\": {",
              json_escape(file).c_str(), c.time_flux, c.time_scnr, c.year0, c.ipx, c.ipy,
              json_escape(c.output_file_full()).c_str());
  for (const auto& f : kPhys) std::printf("\"%s\": %.9g, ", f.name, (double)(c.p.*(f.ptr)));
  std::printf("\"co2_flux\": %.9g, \"p_emi\": [", (double)c.p.co2_flux);
  for (int i = 0; i < 10; ++i) std::printf("%s%.9g", i ? ", " : "", (double)c.p.p_emi[i]);
  std::printf("]}, \"co2_ppm\": [");
  for (size_t i = 0; i < c.co2_ppm.size(); ++i) std::printf("%s%.9g", i ? ", " : "", (double)c.co2_ppm[i]);
  std::printf("]}\n");
}

#define LIB(call)                                                                                   \
  do {                                                                                              \
    const int rc_ = (call);                                                                         \
    if (rc_ != GREB_OK) {                                                                           \
      std::fprintf(stderr, "greb: %s failed (%d): %s\n", #call, rc_, greb_b200_last_error(h));      \
      if (h) greb_b200_destroy(h);                                                                  \
      return 3;                                                                                     \
    }                                                                                               \
  } while (0)

// ---- greb-original (reference src/greb.original.model.f90:138-233, shell :16-60) ---------------------------
// log_exp -> process switches (include/greb_b200.h) and the input / CO2 changes of orig:162-166, 178-179, 225,
// 939-951.  Same table as greb_b200/host.py (ORIGINAL_EXPERIMENTS, original_experiment).
unsigned original_switches(int log_exp) {
  const unsigned nocrcl = GREB_SW_NO_HEAT_CIRCULATION | GREB_SW_NO_VAPOR_CIRCULATION;
  const unsigned ebm = GREB_SW_NO_ICE_ALBEDO | GREB_SW_NO_HYDRO | GREB_SW_NO_DEEP_OCEAN | nocrcl;
  switch (log_exp) {
    case 1: case 2: case 3: case 4: return ebm;
    case 5: return GREB_SW_NO_ICE_ALBEDO | GREB_SW_NO_HYDRO | GREB_SW_NO_DEEP_OCEAN;
    case 6: return GREB_SW_NO_HYDRO | GREB_SW_NO_DEEP_OCEAN;
    case 7: return GREB_SW_NO_VAPOR_CIRCULATION | GREB_SW_NO_DEEP_OCEAN;
    case 8: return GREB_SW_VAPOR_DIFFUSION_ONLY | GREB_SW_NO_DEEP_OCEAN;
    case 9: return GREB_SW_NO_DEEP_OCEAN;
    case 10: case 12: return 0;
    case 11: return GREB_SW_LINEAR_VAPOR_EMISSIVITY | GREB_SW_NO_DEEP_OCEAN;
    case 13: return GREB_SW_NO_HYDRO;
    case 14: return GREB_SW_SST_PLUS_1K | GREB_SW_NO_DEEP_OCEAN;
    case 15: return GREB_SW_SST_PLUS_1K | GREB_SW_NO_DEEP_OCEAN | GREB_SW_NO_HYDRO;
    case 16: return GREB_SW_SST_PLUS_1K | GREB_SW_NO_DEEP_OCEAN | GREB_SW_NO_VAPOR_CIRCULATION;
    default: throw NamelistError("log_exp = " + std::to_string(log_exp) + ": not an experiment of greb.original.model.f90 (1..16)");
  }
}

float a1b_co2(float year) {                       // orig:945-951, fp32 like the reference (`year` is a real)
  float co2 = 680.0f;
  if (year <= 2000.0f) co2 = 310.0f + (60.0f / 50.0f) * (year - 1950.0f);
  if (2000.0f < year && year <= 2050.0f) co2 = 370.0f + (150.0f / 50.0f) * (year - 2000.0f);
  if (2050.0f < year && year <= 2100.0f) co2 = 520.0f + (180.0f / 50.0f) * (year - 2050.0f);
  return co2;
}

struct OriginalConfig {
  int time_flux = 0, time_ctrl = 0, time_scnr = 0, log_exp = 0;   // orig:60: log_exp defaults to 0
};

OriginalConfig original_from_namelist(const std::string& text) {
  auto g = parse_namelist(text);
  OriginalConfig c;
  auto geti = [](Group& grp, const char* k, int dflt) {
    auto it = grp.find(k);
    if (it == grp.end() || it->second.empty() || it->second[0].kind == Value::NONE) return dflt;
    return (int)it->second[0].number(k);
  };
  c.time_flux = geti(g["numerics"], "time_flux", 0);
  c.time_ctrl = geti(g["numerics"], "time_ctrl", 0);
  c.time_scnr = geti(g["numerics"], "time_scnr", 0);
  c.log_exp = geti(g["physics"], "log_exp", 0);
  return c;
}

// `output/control`: 730 records of TF_correct (orig:204-206), the control run's monthly means written over them
// from record 1 (orig:209-215, unit 21)
void write_control(const std::string& path, const std::vector<float>& tf, const std::vector<float>& ctrl) {
  std::vector<float> out(std::max(tf.size(), ctrl.size()), 0.0f);
  std::copy(tf.begin(), tf.end(), out.begin());
  std::copy(ctrl.begin(), ctrl.end(), out.begin());
  make_dirs(path);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error(path + ": cannot open output file");
  const size_t put = std::fwrite(out.data(), sizeof(float), out.size(), f);
  std::fclose(f);
  if (put != out.size()) throw std::runtime_error(path + ": short write");
}

}  // namespace

static int run_original(const std::string& file, const std::string& input_dir, int device, const std::string& arith,
                        bool check) {
  OriginalConfig c;
  unsigned sw = 0;
  try {
    c = original_from_namelist(read_text(file));
    sw = original_switches(c.log_exp);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    return 2;
  }
  const bool a1b = c.log_exp == 12 || c.log_exp == 13;
  const float co2_ctrl = a1b ? 298.0f : 340.0f;                        // orig:178-179
  const int tc = std::max(c.time_ctrl, 0), ts = std::max(c.time_scnr, 0);
  std::vector<float> co2((size_t)tc, co2_ctrl);
  for (int y = 0; y < ts; ++y)                                         // the scenario starts in 1940, orig:219
    co2.push_back(a1b ? a1b_co2((float)(1940 + y)) : (c.log_exp >= 14 ? co2_ctrl : 680.0f));   // orig:225, 943
  if (check) {
    std::printf("{\"namelist\": \"%s\", \"time_flux\": %d, \"time_ctrl\": %d, \"time_scnr\": %d, \"log_exp\": %d, "
                "\"switches\": %u, \"co2_ctrl\": %.9g, \"co2\": [",
                json_escape(file).c_str(), c.time_flux, c.time_ctrl, c.time_scnr, c.log_exp, sw, (double)co2_ctrl);
    for (size_t i = 0; i < co2.size(); ++i) std::printf("%s%.9g", i ? ", " : "", (double)co2[i]);
    std::printf("]}\n");
    return 0;
  }
  const size_t NC = GREB_NCELL, NT = GREB_NSTEP_YR;
  std::vector<float> z_topo, glacier, sw_solar, tclim, qclim, swet, uclim, vclim, mld, cld;
  try {
    z_topo = read_field(input_dir, "topography", NC);
    glacier = read_field(input_dir, "glacier.masks", NC);
    sw_solar = read_field(input_dir, "solar.radiation", NT * GREB_YDIM);
    tclim = read_field(input_dir, "tsurf", NT * NC);
    qclim = read_field(input_dir, "vapor", NT * NC);
    swet = read_field(input_dir, "soil.moisture", NT * NC);
    uclim = read_field(input_dir, "zonal.wind", NT * NC);
    vclim = read_field(input_dir, "meridional.wind", NT * NC);
    mld = read_field(input_dir, "ocean.mld", NT * NC);
    cld = read_field(input_dir, "cloud.cover", NT * NC);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    return 2;
  }
  greb_physics_par p;
  greb_b200_physics_original(&p);
  if (c.log_exp == 1)                                                  // orig:162 constant topography
    for (auto& z : z_topo) z = z > 1.0f ? 1.0f : z;
  if (c.log_exp <= 2) std::fill(cld.begin(), cld.end(), 0.7f);         // orig:163 constant cloud cover
  if (c.log_exp <= 3) std::fill(qclim.begin(), qclim.end(), 0.0052f);  // orig:164 constant water vapour
  if (c.log_exp <= 9 || c.log_exp == 11) std::fill(mld.begin(), mld.end(), p.d_ocean);   // orig:165-166
  p.co2_flux = co2_ctrl;
  std::printf(" %% time flux/control/scenario:  %d %d %d\n", c.time_flux, c.time_ctrl, c.time_scnr);   // shell:59

  greb_b200_t h = nullptr;
  LIB(greb_b200_create(&h, 1, device));
  LIB(greb_b200_set_arithmetic(h, arith == "fast" ? GREB_ARITH_FAST : GREB_ARITH_EXACT));
  LIB(greb_b200_set_forcing(h, z_topo.data(), glacier.data(), sw_solar.data(), tclim.data(), qclim.data(), swet.data(),
                            uclim.data(), vclim.data(), mld.data(), cld.data()));
  const float one = 680.0f;
  LIB(greb_b200_set_member(h, 0, &p, co2.empty() ? &one : co2.data(), co2.empty() ? 1 : (int)co2.size(), 1970));
  LIB(greb_b200_set_switches(h, 0, sw & ~(unsigned)GREB_SW_SST_PLUS_1K));   // orig:226 applies to the scenario only
  LIB(greb_b200_init(h));
  LIB(greb_b200_spinup(h, c.time_flux));                               // orig:201
  std::vector<float> tf(NT * NC);
  LIB(greb_b200_get_fluxcorr(h, 0, 0, tf.data()));                     // orig:204-206
  std::vector<float> ini[4];
  for (int k = 0; k < 4; ++k) {
    ini[k].resize(NC);
    LIB(greb_b200_get_state(h, 0, k, ini[k].data()));
  }
  LIB(greb_b200_reset_scenario(h));
  const size_t year_floats = (size_t)12 * GREB_NVAR_OUT * NC;
  std::vector<float> ctrl((size_t)tc * year_floats), scen((size_t)ts * year_floats), gmc((size_t)tc), gms((size_t)ts);
  LIB(greb_b200_run(h, tc, ctrl.data(), nullptr, 1, gmc.data(), nullptr));        // orig:209-215
  for (int k = 0; k < 4; ++k) LIB(greb_b200_set_state(h, 0, k, ini[k].data()));   // orig:219
  LIB(greb_b200_set_switches(h, 0, sw));
  LIB(greb_b200_run(h, ts, scen.data(), nullptr, 1, gms.data(), nullptr));        // orig:220-231
  for (int y = 0; y < tc; ++y) std::printf("   %12.6f   %12.6f   %12.8f\n", (double)(1970 + y), (double)co2_ctrl, (double)gmc[(size_t)y]);
  for (int y = 0; y < ts; ++y)
    std::printf("   %12.6f   %12.6f   %12.8f\n", (double)(1940 + y), (double)co2[(size_t)(tc + y)], (double)gms[(size_t)y]);
  int rc = 0;
  try {
    write_control("output/control", tf, ctrl);
    write_output("output/scenario", scen.data(), scen.size());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    rc = 2;
  }
  greb_b200_destroy(h);
  return rc;
}

struct Inputs {
  const float *z_topo, *glacier, *sw_solar, *tclim, *qclim, *swet, *uclim, *vclim, *mld, *cld;
};

// one handle = one GPU: spin-up + scenario of `n` members, their output files, their console lines
static int run_block(const RunConfig* cfg, int n, int device, const std::string& arith, const Inputs& in,
                     std::vector<std::string>* lines) {
  const int N = n, tf = cfg[0].time_flux, ts = cfg[0].time_scnr;
  const size_t NC = GREB_NCELL;
  greb_b200_t h = nullptr;
  LIB(greb_b200_create(&h, N, device));
  LIB(greb_b200_set_arithmetic(h, arith == "fast" ? GREB_ARITH_FAST : GREB_ARITH_EXACT));
  LIB(greb_b200_set_forcing(h, in.z_topo, in.glacier, in.sw_solar, in.tclim, in.qclim, in.swet, in.uclim, in.vclim,
                            in.mld, in.cld));
  const float co2_default = 680.0f;
  for (int m = 0; m < N; ++m)
    LIB(greb_b200_set_member(h, m, &cfg[m].p, ts > 0 ? cfg[m].co2_ppm.data() : &co2_default, ts > 0 ? ts : 1,
                             cfg[m].year0));
  LIB(greb_b200_init(h));
  LIB(greb_b200_spinup(h, tf));                                         // f:221
  LIB(greb_b200_reset_scenario(h));                                     // f:226-227
  const size_t year_floats = (size_t)12 * GREB_NVAR_OUT * NC;
  std::vector<std::vector<float>> monthly((size_t)N, std::vector<float>((size_t)(ts > 0 ? ts : 0) * year_floats));
  std::vector<float> year((size_t)N * year_floats), gm((size_t)N);
  static const double days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  for (int y = 0; y < ts; ++y) {                                        // one simulated year per launch
    LIB(greb_b200_run(h, 1, year.data(), nullptr, N, gm.data(), nullptr));
    for (int m = 0; m < N; ++m) {
      const float* rec = year.data() + (size_t)m * year_floats;
      std::memcpy(monthly[(size_t)m].data() + (size_t)y * year_floats, rec, year_floats * sizeof(float));
      // f:954 prints tsmn(ipx,ipy)-273.15, the annual mean of Tsurf at the diagnostic point; the ABI returns
      // monthly means, so the point value is their day-weighted mean
      double s = 0, w = 0;
      for (int mo = 0; mo < 12; ++mo) {
        s += (double)rec[((size_t)mo * GREB_NVAR_OUT + 0) * NC + (size_t)(cfg[m].ipy - 1) * GREB_XDIM + (cfg[m].ipx - 1)] * days[mo];
        w += days[mo];
      }
      const float point = (float)(s / w - 273.15);
      char buf[160];
      std::snprintf(buf, sizeof buf, "   %12.6f   %12.6f   %12.8f   %12.8f\n", (double)(cfg[m].year0 + y),
                    (double)cfg[m].co2_ppm[(size_t)y], (double)gm[(size_t)m], (double)point);
      lines[m].push_back(buf);
    }
  }
  int rc = 0;
  try {
    for (int m = 0; m < N; ++m)
      write_output(cfg[m].output_file_full(), monthly[(size_t)m].data(), monthly[(size_t)m].size());   // opened even for 0 years, f:174
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    rc = 2;
  }
  greb_b200_destroy(h);
  return rc;
}

int main(int argc, char** argv) {
  std::vector<std::string> files;
  std::string input_dir = "input", arith = "exact";
  int device = 0, gpus = 1;
  bool check = false, original = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--check") check = true;
    else if (a == "--original") original = true;
    else if (a == "--input" && i + 1 < argc) input_dir = argv[++i];
    else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "--arith" && i + 1 < argc) arith = argv[++i];
    else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (a == "-h" || a == "--help" || (a.size() > 1 && a[0] == '-' && a[1] == '-')) {
      std::fprintf(stderr, "usage: greb [--check] [--original] [--input DIR] [--device N] [--gpus G] [--arith exact|fast] [namelist ...]\n");
      return a == "-h" || a == "--help" ? 0 : 2;
    } else files.push_back(a);
  }
  if (arith != "exact" && arith != "fast") {
    std::fprintf(stderr, "greb: --arith must be exact or fast\n");
    return 2;
  }
  if (original) return run_original(files.empty() ? "namelist_original" : files[0], input_dir, device, arith, check);
  if (files.empty()) files.push_back("namelist");                       // f:1031-1032
  std::vector<RunConfig> cfg;
  try {
    for (const auto& f : files) cfg.push_back(config_from_namelist(read_text(f)));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    return 2;
  }
  if (check) {
    for (size_t m = 0; m < cfg.size(); ++m) print_check(files[m], cfg[m]);
    return 0;
  }
  const int N = (int)cfg.size(), tf = cfg[0].time_flux, ts = cfg[0].time_scnr;
  for (const auto& c : cfg)
    if (c.time_flux != tf || c.time_scnr != ts) {
      std::fprintf(stderr, "greb: all namelists of one batch must share time_flux and time_scnr\n");
      return 2;
    }
  const size_t NC = GREB_NCELL, NT = GREB_NSTEP_YR;
  std::vector<float> z_topo, glacier, sw_solar, tclim, qclim, swet, uclim, vclim, mld, cld;
  try {                                                                 // f:1018-1027, 1073-1085
    z_topo = read_field(input_dir, "topography", NC);
    glacier = read_field(input_dir, "glacier.masks", NC);
    sw_solar = read_field(input_dir, "solar.radiation", NT * GREB_YDIM);
    tclim = read_field(input_dir, "tsurf", NT * NC);
    qclim = read_field(input_dir, "vapor", NT * NC);
    swet = read_field(input_dir, "soil.moisture", NT * NC);
    uclim = read_field(input_dir, "zonal.wind", NT * NC);
    vclim = read_field(input_dir, "meridional.wind", NT * NC);
    mld = read_field(input_dir, "ocean.mld", NT * NC);
    cld = read_field(input_dir, "cloud.cover", NT * NC);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "greb: %s\n", e.what());
    return 2;
  }
  for (const auto& c : cfg)
    std::printf(" %% diagonstic point lat/lon:  %g %g\n", 3.75 * c.ipy - 90, 3.75 * c.ipx);   // f:1070

  // members in contiguous blocks over the GPUs (heights differ by at most one), one handle and one host thread
  // per GPU; members never interact, so there is no exchange between the blocks
  const int G = std::max(1, std::min(gpus, N));
  const Inputs in{z_topo.data(), glacier.data(), sw_solar.data(), tclim.data(), qclim.data(), swet.data(),
                  uclim.data(), vclim.data(), mld.data(), cld.data()};
  std::vector<std::vector<std::string>> lines((size_t)N);
  std::vector<int> rcs((size_t)G, 0);
  std::vector<std::thread> threads;
  for (int g = 0; g < G; ++g) {
    const int m0 = (int)((long)N * g / G), m1 = (int)((long)N * (g + 1) / G);
    threads.emplace_back([&, g, m0, m1] {
      rcs[(size_t)g] = run_block(cfg.data() + m0, m1 - m0, device + g, arith, in, lines.data() + m0);
    });
  }
  for (auto& t : threads) t.join();
  for (int y = 0; y < ts; ++y)                                          // the console lines, year by year
    for (int m = 0; m < N; ++m)
      if ((size_t)y < lines[(size_t)m].size()) std::fputs(lines[(size_t)m][(size_t)y].c_str(), stdout);
  for (int rc : rcs)
    if (rc != 0) return rc;
  return 0;
}
