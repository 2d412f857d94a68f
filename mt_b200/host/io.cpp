/*
 * io.cpp — configuration table and file formats of the drop-in host.
 *
 * Behaviour contracts (not code) taken from the reference:
 *   config  src/configreader.cpp:24-76 (parse), :99-131 (lookup / DIE), :134-231 (typed getters),
 *           :260-268 (token trim + mask), :329-363 (mask / replace)
 *   PDB     src/pdbio.cpp:36-80 (read), :171-213 (fixed columns), :276-371 (write / append)
 *   DCD     src/dcdio.cpp:16-39 (header fields), :98-150 (header layout), :179-203 (frame)
 *   XYZ     src/xyzio.cpp:16-44 (read), :73-92 (write)
 */
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include "mt_host.hpp"

namespace mt {

void die(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Fatal(buf);
}

// =================================================================== ParamTable
static bool skipped_line(const std::string &l)
{
    return l.empty() || l[0] == '#' || l[0] == ' ' || l[0] == '\n' || l[0] == '\t';
}

void ParamTable::parse(const std::string &filename, const std::vector<std::string> &overrides)
{
    FILE *f = fopen(filename.c_str(), "r");
    if (!f) die("Opening file '%s'", filename.c_str());
    if (!quiet) printf("Parsing '%s' parameters file...\n", filename.c_str());
    items_.clear();
    char buf[4096];
    while (fgets(buf, sizeof buf, f)) {
        std::string line(buf);
        if (skipped_line(line)) continue;
        // name: up to the first blank/tab; value: the rest up to '#' or end of line, untrimmed
        size_t e = line.find_first_of(" \t");
        std::string name = line.substr(0, e), value;
        if (e != std::string::npos) {
            size_t b = e + 1;
            while (b < line.size() && (line[b] == '#' || line[b] == '\n')) b++; // strtok skips leading delimiters
            size_t stop = line.find_first_of("#\n", b);
            value = line.substr(b, stop == std::string::npos ? std::string::npos : stop - b);
        }
        if (!name.empty() && name.back() == '\n') name.pop_back();
        if (!quiet) printf("%s\t%s\n", name.c_str(), value.c_str());
        items_.emplace_back(name, value);
    }
    fclose(f);
    for (const std::string &o : overrides) {
        size_t eq = o.find('=');
        if (eq == std::string::npos) continue;
        if (!quiet) printf("Enforced via argv: %s\t=%s\n", o.substr(0, eq).c_str(), o.substr(eq + 1).c_str());
        set(o.substr(0, eq), o.substr(eq + 1), true);
    }
}

void ParamTable::set(const std::string &name, const std::string &value, bool force)
{
    for (auto &it : items_)
        if (it.first == name) {
            it.second = value;
            return;
        }
    if (force) items_.emplace_back(name, value);
}

bool ParamTable::raw(const std::string &name, std::string &value) const
{
    for (const auto &it : items_)
        if (it.first == name) {
            value = it.second;
            return true;
        }
    return false;
}

bool ParamTable::lookup(const std::string &name, const std::string *def, std::string &out)
{
    if (raw(name, out)) {
        if (!quiet) printf("'%s' = '%s'\n", name.c_str(), out.c_str());
        return true;
    }
    if (!def) die("Parameter '%s' should be specified in a configuration file.", name.c_str());
    out = *def;
    if (!quiet) printf("Using default value for parameter %s: '%s' = '%s'\n", name.c_str(), name.c_str(), def->c_str());
    return true;
}

static std::string first_token(const std::string &v)
{
    size_t b = v.find_first_not_of(" \t");
    if (b == std::string::npos) return "";
    size_t e = v.find_first_of(" \t", b);
    return v.substr(b, e == std::string::npos ? std::string::npos : e - b);
}

static std::string replace_first(const std::string &s, const std::string &what, const std::string &with)
{
    size_t p = s.find(what);
    if (p == std::string::npos) return s;
    return s.substr(0, p) + with + s.substr(p + what.size());
}

std::string ParamTable::apply_mask(const std::string &token) const
{
    std::string r = token;
    for (const auto &it : items_) r = replace_first(r, "<" + it.first + ">", first_token(it.second));
    return r;
}

std::string ParamTable::masked(const std::string &name)
{
    std::string v;
    lookup(name, nullptr, v);
    return apply_mask(first_token(v));
}
std::string ParamTable::masked(const std::string &name, const std::string &def)
{
    std::string v;
    lookup(name, &def, v);
    return apply_mask(first_token(v));
}
std::string ParamTable::masked_replace(const std::string &name, const std::string &replacement, const std::string &what)
{
    return replace_first(masked(name), what, replacement);
}

int ParamTable::integer(const std::string &name)
{
    std::string v = masked(name);
    int r = atoi(v.c_str());
    if (r == 0 && v != "0") die("ERROR: Wrong value of %s in a configuration file ('%s'). Should be integer.", name.c_str(), v.c_str());
    return r;
}
int ParamTable::integer(const std::string &name, int def)
{
    std::string v = masked(name, std::to_string(def));
    int r = atoi(v.c_str());
    if (r == 0 && v != "0") die("ERROR: Wrong value of %s in a configuration file ('%s'). Should be integer.", name.c_str(), v.c_str());
    return r;
}
long long ParamTable::long_integer(const std::string &name, long def)
{
    std::string v = masked(name, std::to_string(def));
    long long r = atol(v.c_str());
    if (r == 0 && v != "0")
        die("ERROR: Wrong value of %s in a configuration file ('%s'). Should be long integer.", name.c_str(), v.c_str());
    return r;
}
static float parse_float(const std::string &name, const std::string &v)
{
    float r = (float)atof(v.c_str());
    if (r == 0.0 && v != "0" && v != "0.0" && v != "0.0f" && v != "0.000000")
        die("ERROR: Wrong value of %s in a configuration file ('%s'). Should be float.", name.c_str(), v.c_str());
    return r;
}
float ParamTable::real(const std::string &name) { return parse_float(name, masked(name)); }
float ParamTable::real(const std::string &name, float def)
{
    char d[64];
    snprintf(d, sizeof d, "%f", def);
    return parse_float(name, masked(name, d));
}
int ParamTable::yesno(const std::string &name, int def, bool allow_default)
{
    std::string v = allow_default ? masked(name, def ? "YES" : "NO") : masked(name);
    static const char *yes[] = {"YES", "Yes", "yes", "Y", "y", "ON", "On", "on", "TRUE", "True", "true"};
    for (const char *y : yes)
        if (v == y) return 1;
    return 0;
}

// =================================================================== PDB
static std::string field(const std::string &l, size_t pos, size_t len)
{
    if (pos >= l.size()) return "";
    return l.substr(pos, len);
}

void read_pdb(const std::string &filename, PDB &pdb, bool quiet)
{
    if (!quiet) printf("Reading %s.\n", filename.c_str());
    FILE *f = fopen(filename.c_str(), "r");
    if (!f) {
        perror(filename.c_str());
        die("cannot read PDB '%s'", filename.c_str());
    }
    pdb.atoms.clear();
    char buf[80]; // the reference reads 79-character chunks (BUF_SIZE 80)
    while (fgets(buf, sizeof buf, f)) {
        // record name must be exactly "ATOM" followed by a blank (strtok(buffer," ") == "ATOM")
        if (strncmp(buf, "ATOM", 4) != 0 || (buf[4] != ' ' && buf[4] != '\0')) continue;
        std::string l(buf);
        PDBAtom a;
        a.id = atoi(field(l, 6, 5).c_str());
        std::string nm = first_token(field(l, 12, 4));
        // strtok(atomName, " ") only splits on blanks
        {
            std::string raw = field(l, 12, 4);
            size_t b = raw.find_first_not_of(' ');
            size_t e = b == std::string::npos ? b : raw.find(' ', b);
            nm = b == std::string::npos ? "" : raw.substr(b, e == std::string::npos ? std::string::npos : e - b);
        }
        strncpy(a.name, nm.c_str(), 4);
        a.altLoc = l.size() > 16 ? l[16] : ' ';
        strncpy(a.resName, field(l, 17, 3).c_str(), 3);
        a.chain = l.size() > 21 ? l[21] : ' ';
        a.resid = atoi(field(l, 22, 4).c_str());
        a.x = atof(field(l, 30, 8).c_str());
        a.y = atof(field(l, 38, 8).c_str());
        a.z = atof(field(l, 46, 8).c_str());
        a.occupancy = atof(field(l, 54, 6).c_str());
        a.beta = atof(field(l, 60, 6).c_str());
        pdb.atoms.push_back(a);
    }
    fclose(f);
    if (!quiet) {
        printf("Found %d atoms.\n", (int)pdb.atoms.size());
        printf("Done reading '%s'.\n", filename.c_str());
    }
}

// ATOM lines, byte-identical to the reference's
//   fprintf("ATOM  %5d %-4s%c%3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n")      (pdbio.cpp:291-345; %5x from 100000 atoms)
// but formatted with std::to_chars (correctly rounded like printf, ~10x faster: the hydrolysis dump prints
// Ntot*Ntr lines every 10 strides) into one buffer and written with a single fwrite.
static char *put_fixed(char *o, double v, int width, int prec)
{
    char tmp[64];
    auto res = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::fixed, prec);
    const int len = (int)(res.ptr - tmp);
    for (int k = len; k < width; k++) *o++ = ' ';
    memcpy(o, tmp, len);
    return o + len;
}
static char *put_int(char *o, long v, int width, int base = 10)
{
    char tmp[32];
    auto res = std::to_chars(tmp, tmp + sizeof tmp, v, base);
    const int len = (int)(res.ptr - tmp);
    for (int k = len; k < width; k++) *o++ = ' ';
    memcpy(o, tmp, len);
    return o + len;
}
static void print_atoms(FILE *f, const PDB &pdb)
{
    const bool hex = pdb.atoms.size() >= 100000;
    std::vector<char> buf;
    buf.resize(pdb.atoms.size() * 96 + 16);
    char *o = buf.data();
    for (size_t i = 0; i < pdb.atoms.size(); i++) {
        const PDBAtom &a = pdb.atoms[i];
        if (!std::isfinite(a.x) || !std::isfinite(a.y) || !std::isfinite(a.z) || fabs(a.x) > 1e15 || fabs(a.y) > 1e15 || fabs(a.z) > 1e15) {
            // keep printf's spelling of nan/inf and of absurdly wide numbers
            o += snprintf(o, 200, hex ? "ATOM  %5x %-4s%c%3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n"
                                      : "ATOM  %5d %-4s%c%3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n",
                          (int)i + 1, a.name, a.altLoc, a.resName, a.chain, a.resid, a.x, a.y, a.z, a.occupancy, a.beta);
        } else {
            memcpy(o, "ATOM  ", 6);
            o += 6;
            o = put_int(o, (long)i + 1, 5, hex ? 16 : 10);
            *o++ = ' ';
            int nl = (int)strnlen(a.name, 4);
            memcpy(o, a.name, nl);
            o += nl;
            for (int k = nl; k < 4; k++) *o++ = ' ';
            *o++ = a.altLoc;
            int rl = (int)strnlen(a.resName, 3);
            for (int k = rl; k < 3; k++) *o++ = ' ';
            memcpy(o, a.resName, rl);
            o += rl;
            *o++ = ' ';
            *o++ = a.chain;
            o = put_int(o, a.resid, 4);
            memcpy(o, "    ", 4);
            o += 4;
            o = put_fixed(o, a.x, 8, 3);
            o = put_fixed(o, a.y, 8, 3);
            o = put_fixed(o, a.z, 8, 3);
            o = put_fixed(o, a.occupancy, 6, 2);
            o = put_fixed(o, a.beta, 6, 2);
            *o++ = '\n';
        }
        if ((size_t)(o - buf.data()) + 256 > buf.size()) { // lines longer than expected: flush
            fwrite(buf.data(), 1, o - buf.data(), f);
            o = buf.data();
        }
    }
    fwrite(buf.data(), 1, o - buf.data(), f);
}
void write_pdb(const std::string &filename, const PDB &pdb, bool quiet)
{
    if (!quiet) printf("Saving PDB '%s'...\n", filename.c_str());
    FILE *f = fopen(filename.c_str(), "w");
    if (!f) die("Opening file '%s'", filename.c_str());
    print_atoms(f, pdb);
    fprintf(f, "END");
    fclose(f);
    if (!quiet) printf("Done saving PDB.\n");
}
void append_pdb(const std::string &filename, const PDB &pdb, bool quiet)
{
    if (!quiet) printf("Appending PDB '%s'...\n", filename.c_str());
    FILE *f = fopen(filename.c_str(), "a");
    if (!f) die("Opening file '%s'", filename.c_str());
    print_atoms(f, pdb);
    fprintf(f, "END\n");
    fclose(f);
    if (!quiet) printf("Done appending PDB.\n");
}

// =================================================================== DCD (CHARMM/NAMD, no unit cell)
DCDHeader make_dcd_header(int n_atoms, int frame_count, int first_frame, float timestep, int dcd_freq)
{
    DCDHeader h;
    h.n_atoms = n_atoms;
    h.delta = timestep;
    h.nfile = frame_count;
    h.npriv = first_frame;
    h.nsavc = dcd_freq;
    h.remark1 = "REMARKS CREATED BY dcdio.c";
    time_t raw;
    time(&raw);
    h.remark2 = std::string("REMARKS DATE: ") + asctime(localtime(&raw));
    return h;
}

static void put_i32(FILE *f, int v) { fwrite(&v, 4, 1, f); }
static void put_remark(FILE *f, const std::string &s)
{
    char t[80];
    memset(t, 0, sizeof t);
    memcpy(t, s.data(), s.size() < 80 ? s.size() : 79);
    fwrite(t, 80, 1, f);
}

void dcd_write_header(FILE *f, const DCDHeader &h)
{
    put_i32(f, 84);
    fwrite("CORD", 4, 1, f);
    put_i32(f, h.nfile);
    put_i32(f, h.npriv);
    put_i32(f, h.nsavc);
    put_i32(f, h.npriv - h.nsavc);
    for (int i = 0; i < 5; i++) put_i32(f, 0);
    fwrite(&h.delta, 4, 1, f);
    put_i32(f, 0); // no unit cell
    for (int i = 0; i < 8; i++) put_i32(f, 0);
    put_i32(f, 24);
    put_i32(f, 84);
    put_i32(f, 164);
    put_i32(f, 2);
    put_remark(f, h.remark1);
    put_remark(f, h.remark2);
    put_i32(f, 164);
    put_i32(f, 4);
    put_i32(f, h.n_atoms);
    put_i32(f, 4);
}

void dcd_write_frame(FILE *f, int n, const float *x, const float *y, const float *z)
{
    const float *c[3] = {x, y, z};
    for (int k = 0; k < 3; k++) {
        put_i32(f, n * 4);
        fwrite(c[k], 4, (size_t)n, f);
        put_i32(f, n * 4);
    }
}

static bool get_i32(FILE *f, int &v) { return fread(&v, 4, 1, f) == 1; }

bool dcd_read_header(FILE *f, DCDHeader &h)
{
    int v;
    char cord[4];
    if (!get_i32(f, v) || v != 84) return false;
    if (fread(cord, 4, 1, f) != 1 || strncmp(cord, "CORD", 4) != 0) return false;
    get_i32(f, h.nfile);
    get_i32(f, h.npriv);
    get_i32(f, h.nsavc);
    for (int i = 0; i < 6; i++) get_i32(f, v);
    if (fread(&h.delta, 4, 1, f) != 1) return false;
    for (int i = 0; i < 9; i++) get_i32(f, v);
    for (int i = 0; i < 4; i++) get_i32(f, v); // 24, 84, 164, 2
    char t[81] = {0};
    if (fread(t, 80, 1, f) != 1) return false;
    h.remark1 = t;
    if (fread(t, 80, 1, f) != 1) return false;
    h.remark2 = t;
    get_i32(f, v);
    get_i32(f, v);
    if (!get_i32(f, h.n_atoms)) return false;
    return get_i32(f, v);
}

bool dcd_read_frame(FILE *f, int n, float *x, float *y, float *z)
{
    float *c[3] = {x, y, z};
    for (int k = 0; k < 3; k++) {
        int v;
        if (!get_i32(f, v)) return false;
        if (fread(c[k], 4, (size_t)n, f) != (size_t)n) return false;
        if (!get_i32(f, v)) return false;
    }
    return true;
}

// =================================================================== XYZ
void read_xyz(const std::string &filename, std::vector<XYZAtom> &atoms, bool quiet)
{
    if (!quiet) printf("Reading %s.\n", filename.c_str());
    FILE *f = fopen(filename.c_str(), "r");
    if (!f) die("Opening file '%s'", filename.c_str());
    char buf[256];
    if (!fgets(buf, sizeof buf, f)) die("Error reading '%s'", filename.c_str());
    int n = atoi(buf);
    if (!fgets(buf, sizeof buf, f)) die("Error reading '%s'", filename.c_str());
    atoms.resize(n > 0 ? n : 0);
    for (int i = 0; i < n; i++) {
        if (!fgets(buf, sizeof buf, f)) die("Error reading '%s': %d atoms announced, %d found", filename.c_str(), n, i);
        char *tok = strtok(buf, " \t\r\n");
        atoms[i].name = tok ? tok[0] : ' ';
        tok = strtok(NULL, " \t\r\n");
        atoms[i].x = tok ? atof(tok) : 0;
        tok = strtok(NULL, " \t\r\n");
        atoms[i].y = tok ? atof(tok) : 0;
        tok = strtok(NULL, " \t\r\n");
        atoms[i].z = tok ? atof(tok) : 0;
    }
    fclose(f);
    if (!quiet) printf("Done reading '%s'.\n", filename.c_str());
}

void write_xyz(const std::string &filename, const std::vector<XYZAtom> &atoms, bool quiet)
{
    if (!quiet) printf("Writing %s.\n", filename.c_str());
    FILE *f = fopen(filename.c_str(), "w");
    if (!f) die("Opening file '%s'", filename.c_str());
    fprintf(f, "%d\n", (int)atoms.size());
    fprintf(f, "Created by 'xyzio.cpp'\n");
    for (const XYZAtom &a : atoms) fprintf(f, "%-*c%*f%*f%*f\n", 16, a.name, 16, a.x, 16, a.y, 16, a.z);
    fclose(f);
    if (!quiet) printf("Done writing '%s'.\n", filename.c_str());
}

} // namespace mt
