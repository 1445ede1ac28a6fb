// cuCLARK / cuCLARK-l — command line of the B200 build.
//
// Same flags, targets file, database file names, result CSV and stdout/stderr lines as the
// reference executables (src/main.cc:74-320, CuCLARK ctor src/CuCLARK_hh.hh:221-310,
// run/runSimple :383-573, getdbName :580-591, getTargetsData :1795-1906), so that
// classify_metagenome.sh and friends keep working. Everything between "input file" and
// "result CSV" is one call into libcuclark_b200.so (cuclark_classify_file_multi): the host
// does no per-read work. The variant is chosen at compile time as in the reference
// (header swap there, -DCUCLARK_LIGHT here): cuCLARK = HTSIZE 1610612741, default k 31;
// cuCLARK-l = HTSIZE 57777779, k forced to 27.
//
// Differences, all in the direction of "less work for the caller":
//  * -n / -b are accepted and echoed but only size the pipeline (chunks in flight);
//  * -d N (default: every GPU, as in the reference): a table that fits one GPU is loaded once, replicated over NVLink
//    and the reads are dealt over the GPUs; a larger one is partitioned by bucket range as in the reference and the
//    GPUs exchange k-mers and labels (csrc/route.cu);
//  * a missing database is built on the GPU (cuclark_build_database) from FASTA targets, byte-identical
//    to the reference's files; --tsk, spectrum/FASTQ targets and the 3rd targets column are not supported.
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cuclark_b200.h"

#define CLI_VERSION "1.1"
#define MAXK 32
#define SFACTORMAX 30
#ifdef CUCLARK_LIGHT
static const uint64_t HTSIZE = CUCLARK_HTSIZE_LIGHT;
#else
static const uint64_t HTSIZE = CUCLARK_HTSIZE_FULL;
#endif

using std::cerr;
using std::cout;
using std::endl;
using std::string;
using std::vector;

static void print_usage() {
    cout << "\ncuCLARK (B200 build) -- classification of reads against a database of target-specific k-mers\n\n"
         << "./cuCLARK -k <kmerSize> -T <fileTargets> -D <directoryDB/> -O <fileObjects> -R <fileResults> ...\n\n"
         << "-k <kmerSize>         k-mer length, integer in [2,32] (default 31; cuCLARK-l always uses 27)\n"
         << "-t <minFreqTarget>    minimum k-mer frequency in targets (part of the database name)\n"
         << "-T <fileTargets>      targets definition: one '<file> <label>' per line\n"
         << "-D <directoryDB/>     directory holding the database files\n"
         << "-O <fileObjects>      FASTA/FASTQ file with the reads (or a list of such files)\n"
         << "-P <file1> <file2>    paired-end reads (FASTQ)\n"
         << "-R <fileResults>      results are written to <fileResults>.csv (or a list of names)\n"
         << "-n <numberofthreads>  host threads (chunks in flight per GPU)\n"
         << "-b <numberofbatches>  accepted for compatibility\n"
         << "-d <numberofdevices>  number of GPUs (default: all; table replicated if it fits one GPU, partitioned otherwise)\n"
         << "-g <iteration>        gap of the cuCLARK-l database (part of its name; default 4)\n"
         << "-s <factor>           sampling factor when loading the database (cuCLARK only)\n"
         << "--tsk                 accepted for compatibility\n"
         << "--extended            one hit-count column per target in the results\n"
         << "--cache               keep the device-layout table next to the database (<db>.b200) and reuse it\n"
         << "--help, --version\n"
         << endl;
}

static bool valid_file(const char* p) {
    FILE* f = fopen(p, "r");
    if (!f) return false;
    fclose(f);
    return true;
}

// whitespace-separated tokens of a line (src/file.cc:64-87 also splits on ',')
static vector<string> tokens(const string& line, const char* seps, size_t max_n) {
    vector<string> out;
    size_t t = 0;
    while (t < line.size() && out.size() < max_n) {
        while (t < line.size() && strchr(seps, line[t])) t++;
        string v;
        while (t < line.size() && !strchr(seps, line[t])) v.push_back(line[t++]);
        if (!v.empty()) out.push_back(v);
    }
    return out;
}

static bool get_line(FILE* f, string& line) {
    char* buf = nullptr;
    size_t cap = 0;
    const ssize_t n = getline(&buf, &cap, f);
    if (n < 0) { free(buf); line.clear(); return false; }
    line.assign(buf, (size_t)n);
    free(buf);
    if (!line.empty() && line.back() == '\n') line.pop_back();
    return true;
}

struct Cli {
    size_t k = 31, cpu = 1, iter_kmers = 0, batches = 1, devices = 0;
    unsigned min_t = 0, sfactor = 1;
    bool light = false, tsk = false, ext = false, cache = false;
    string targets, folder, results;
    const char *objects = nullptr, *objects2 = nullptr;
    vector<string> labels, labels_c, names;     // names[0] = "NA"
    vector<string> target_files;                // one per line of the targets file
    vector<uint16_t> target_label;              // its label index
    vector<cuclark_db*> dbs;
    size_t n_objects = 0;
};

static string db_name(const Cli& c) {
    // src/CuCLARK_hh.hh:580-591 (the folder already ends in '/', so the name holds "//" as the reference's does)
    char buf[4096];
    const size_t n_labels = c.labels.size() + c.labels_c.size();
    if (c.light)
        snprintf(buf, sizeof buf, "%s/db_central_k%lu_t%lu_s%lu_m%lu_light_%lu.tsk", c.folder.c_str(), c.k, n_labels,
                 (size_t)HTSIZE, (size_t)c.min_t, c.iter_kmers);
    else
        snprintf(buf, sizeof buf, "%s/db_central_k%lu_t%lu_s%lu_m%lu.tsk", c.folder.c_str(), c.k, n_labels,
                 (size_t)HTSIZE, (size_t)c.min_t);
    return buf;
}

// src/CuCLARK_hh.hh:1795-1906: every target file must exist; labels in order of first appearance
static void read_targets(Cli& c) {
    FILE* f = fopen(c.targets.c_str(), "r");
    if (!f) { cerr << "Failed to open targets data in file: " << c.targets << endl; exit(-1); }
    string line;
    while (get_line(f, line)) {
        const vector<string> e = tokens(line, " \t\n\r,", 3);
        if (e.empty()) continue;
        if (!valid_file(e[0].c_str())) {
            cerr << "Failed to open file: " << e[0] << " defined in " << c.targets << endl;
            exit(-1);
        }
        if (e.size() < 2) { cerr << " Missing label for " << e[0] << endl; exit(-1); }
        if (std::find(c.labels.begin(), c.labels.end(), e[1]) == c.labels.end()) c.labels.push_back(e[1]);
        c.target_files.push_back(e[0]);
        c.target_label.push_back((uint16_t)(std::find(c.labels.begin(), c.labels.end(), e[1]) - c.labels.begin()));
        if (e.size() > 2 && std::find(c.labels_c.begin(), c.labels_c.end(), e[2]) == c.labels_c.end())
            c.labels_c.push_back(e[2]);
    }
    fclose(f);
    c.names.push_back("NA");
    for (auto& l : c.labels) c.names.push_back(l);
    for (auto& l : c.labels_c) c.names.push_back(l);
}

static void banner(const Cli& c) {
    cerr << "CuCLARK version " << CLI_VERSION << " (Copyright 2016 Robin Kobus, rkobus@students.uni-mainz.de)" << endl;
    cerr << "Based on CLARK version 1.1.3 (UCR CS&E. Copyright 2013-2016 Rachid Ounit, rouni001@cs.ucr.edu) " << endl;
    cerr << "B200 build: libcuclark_b200 " << cuclark_version() << endl;
    if (c.min_t > 0) cerr << "Minimum k-mers occurences in Targets is set to " << c.min_t << endl;
    if (c.tsk) cerr << "Creation of targets specific k-mers files requested " << endl;
    if (c.light) cerr << "Using light database in RAM (" << c.iter_kmers << ")" << endl;
    if (c.sfactor > 2) cerr << "Sampling factor is " << c.sfactor << endl;
}

[[noreturn]] static void die_lib(const char* what) {
    cerr << "CUERR '" << cuclark_last_error() << "' (" << what << ")" << endl;      // src/CuClarkDB.cu:45-63
    exit(1);
}

// makeSpecificTargetSets (src/CuCLARK_hh.hh:691-1112) on the device: runs when the database files are absent
static void build_database(Cli& c, const string& base) {
    cerr << "Starting the creation of the database of targets specific " << c.k << "-mers from input files..." << endl;
    if (!c.labels_c.empty()) {
        cerr << "A third column (centromere labels) in the targets file is not supported by the device builder." << endl;
        exit(-1);
    }
    cuclark_build_opts o;
    memset(&o, 0, sizeof o);
    o.k = (int)c.k;
    o.htsize = HTSIZE;
    o.light_gap = c.light ? (int)c.iter_kmers : 0;
    o.min_count = c.min_t;
    vector<const char*> files;
    for (auto& f : c.target_files) files.push_back(f.c_str());
    cuclark_build_stats st;
    const int rc = cuclark_build_database(&o, files.data(), c.target_label.data(), files.size(), base.c_str(), &st);
    if (rc == CUCLARK_ERR_NO_DEVICE) { cerr << "Not enough CUDA devices found: " << cuclark_last_error() << endl; exit(1); }
    if (rc) die_lib("cuclark_build_database");
    cerr << st.n_nucleotides << " nt read in total." << endl;
    cerr << "Removal of common k-mers done: " << st.n_kmers_kept << " specific " << c.k << "-mers found." << endl;
    cerr << (c.light ? "Creating light database in disk..." : "Creating database in disk...") << endl;
    cerr << st.n_kmers_kept << " " << c.k << "-mers successfully stored in database." << endl;
}

static cuclark_db* create_handle(const Cli& c, int device, int shard, int n_shards) {
    cuclark_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.k = (int)c.k;
    cfg.htsize = HTSIZE;
    cfg.n_targets = (int)(c.names.size() - 1);
    cfg.device = device;
    cfg.shard_index = shard;
    cfg.shard_count = n_shards;
    cuclark_db* db = nullptr;
    const int rc = cuclark_create(&cfg, &db);
    if (rc == CUCLARK_ERR_NO_DEVICE) {
        cerr << "Not enough CUDA devices found: " << cuclark_last_error() << endl;      // src/CuClarkDB.cu:109-118
        exit(1);
    }
    if (rc) die_lib("cuclark_create");
    return db;
}

// .sz/.ky/.lb (or the table cache) into one handle
static void load_into(Cli& c, cuclark_db* db, const string& base, bool announce) {
    // --cache: <base>[.s<f>].b200 holds the table in its device layout; written after the first
    // load from .sz/.ky/.lb, streamed back (no rebuild) while it still matches those files
    const string cache = base + (c.sfactor > 1 ? ".s" + std::to_string(c.sfactor) : string()) + ".b200";
    if (c.cache && valid_file(cache.c_str())) {
        const int rc = cuclark_load_table(db, cache.c_str(), base.c_str(), (int)c.sfactor);
        if (rc == CUCLARK_OK) { if (announce) cerr << "Table cache " << cache << " loaded." << endl; return; }
        if (rc == CUCLARK_ERR_FORMAT || rc == CUCLARK_ERR_IO) cerr << "Ignoring table cache: " << cuclark_last_error() << endl;
        else die_lib("cuclark_load_table");
    }
    const int rc = cuclark_load_db_files(db, base.c_str(), (int)c.sfactor);
    if (rc == CUCLARK_ERR_IO) { cerr << cuclark_last_error() << endl << "Failed to find the database." << endl; exit(-1); }
    if (rc) die_lib("cuclark_load_db_files");
    if (c.cache && announce) {
        if (cuclark_save_table(db, cache.c_str()) == CUCLARK_OK) cerr << "Table cache " << cache << " written." << endl;
        else cerr << "Table cache not written: " << cuclark_last_error() << endl;
    }
}

// CuClarkDB ctor + read (src/CuClarkDB.cu:85-208, 462-808). `-d` absent = every GPU found (src/main.cc:104,
// src/CuClarkDB.cu:108-138). The reference ALWAYS partitions the table over the devices (:546-574). Here a table that
// fits one device is loaded once and replicated over NVLink, the reads are dealt over the devices (no exchange at
// all); a table that does not fit one device is partitioned by bucket range as in the reference, and the devices
// exchange k-mers and labels (csrc/route.cu). CUCLARK_PARTITION_TABLE=1 forces the second mode.
static void load_database(Cli& c) {
    const string base = db_name(c);
    bool present = true;
    for (const char* ext : {".sz", ".ky", ".lb"}) present = present && valid_file((base + ext).c_str());
    if (!present) build_database(c, base);
    int n_avail = 0;
    uint64_t free0 = 0, total0 = 0;
    if (cuclark_device_info(0, &n_avail, &free0, &total0) != CUCLARK_OK) {
        cerr << "Not enough CUDA devices found: " << cuclark_last_error() << endl;          // src/CuClarkDB.cu:109-118
        exit(1);
    }
    int n_dev = c.devices ? (int)c.devices : n_avail;
    if (n_dev > n_avail) {
        cerr << "Not enough CUDA devices found: " << n_avail << " of " << n_dev << " requested." << endl;   // :126-131
        exit(1);
    }
    if (n_dev > 16) n_dev = 16;
    // does the table fit one device? entries = bytes of .ky / key width (src/main.cc:278-316)
    bool partition = getenv("CUCLARK_PARTITION_TABLE") != nullptr && n_dev > 1;
    {
        const size_t t_b = (size_t)(log((double)HTSIZE) / log(4.0));
        const int kb = c.k <= t_b + 8 ? 2 : c.k <= t_b + 16 ? 4 : 8;
        FILE* f = fopen((base + ".ky").c_str(), "rb");
        uint64_t n_entries = 0;
        if (f) { fseeko(f, 0, SEEK_END); n_entries = (uint64_t)ftello(f) / kb; fclose(f); }
        cuclark_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.k = (int)c.k; cfg.htsize = HTSIZE; cfg.n_targets = (int)(c.names.size() - 1); cfg.shard_count = 1;
        cuclark_table_plan plan;
        if (cuclark_plan_table(&cfg, n_entries / (c.sfactor > 1 ? c.sfactor : 1), &plan) == CUCLARK_OK) {
            const double need = (double)plan.home_bytes * 1.10 + 3e9;      // + overflow table, build scratch, read buffers
            if (need > (double)free0) {
                if (n_dev < 2) {
                    cerr << "The database needs about " << (size_t)(need / 1e9) << " GB of device memory, the device has "
                         << (size_t)(free0 / 1e9) << " GB free: use more devices (-d)." << endl;
                    exit(1);
                }
                partition = true;
            }
        }
    }
    cerr << "Loading database [" << base << ".*] (s=" << c.sfactor << ")..." << endl;
    if (!partition) {
        c.dbs.push_back(create_handle(c, 0, 0, 1));
        load_into(c, c.dbs[0], base, true);
        for (int d = 1; d < n_dev; d++) {                    // replicas: device-to-device copies of the built table
            cuclark_db* db = create_handle(c, d, 0, 1);
            if (cuclark_clone_table(c.dbs[0], db) != CUCLARK_OK) die_lib("cuclark_clone_table");
            c.dbs.push_back(db);
        }
        if (n_dev > 1) cerr << "Using " << n_dev << " devices: table replicated, reads partitioned." << endl;
        return;
    }
    // table-partitioned: shard d of n_dev on device d, loaded concurrently (the files are read through the page cache)
    for (int d = 0; d < n_dev; d++) c.dbs.push_back(create_handle(c, d, d, n_dev));
    vector<int> rcs(n_dev, 0);
    vector<string> errs(n_dev);
    vector<std::thread> th;
    for (int d = 0; d < n_dev; d++)
        th.emplace_back([&, d] {
            rcs[d] = cuclark_load_db_files(c.dbs[d], base.c_str(), (int)c.sfactor);
            if (rcs[d]) errs[d] = cuclark_last_error();
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; d++) {
        if (rcs[d] == CUCLARK_ERR_IO) { cerr << errs[d] << endl << "Failed to find the database." << endl; exit(-1); }
        if (rcs[d]) { cerr << "CUERR '" << errs[d] << "' (cuclark_load_db_files, device " << d << ")" << endl; exit(1); }
    }
    cerr << "Using " << n_dev << " devices: table partitioned by bucket range, k-mers routed to their shard." << endl;
}

// src/CuCLARK_hh.hh:512-573
static int csv_file_sink(void* user, const char* data, size_t n, uint64_t offset) {
    const int fd = (int)(intptr_t)user;
    while (n) {
        const ssize_t w = pwrite(fd, data, n, (off_t)offset);
        if (w <= 0) return -1;
        data += w; n -= (size_t)w; offset += (uint64_t)w;
    }
    return 0;
}

// `memory`: the reads are already in memory (joined mates) instead of in the file `objects`
static void run_simple(Cli& c, const char* objects, const char* result, bool paired, const vector<char>* memory = nullptr) {
    const string csv = string(result) + ".csv";
    vector<const char*> name_ptrs;
    for (size_t t = 1; t < c.names.size(); t++) name_ptrs.push_back(c.names[t].c_str());
    cuclark_text_opts o;
    memset(&o, 0, sizeof o);
    o.paired = paired;
    o.extended = c.ext;
    o.target_names = name_ptrs.data();
    o.n_slots = (int)std::max<size_t>(4, std::min<size_t>(c.cpu, 8));      // chunks in flight per GPU (16 slots: 0.22 -> 0.3 s per 10 M reads, slot allocation)
    if (const char* e = getenv("CUCLARK_CHUNK_MB")) o.chunk_bytes = (size_t)atol(e) << 20;
    cuclark_text_stats st;
    struct timeval t0, t1;
    gettimeofday(&t0, nullptr);
    cerr << (c.ext ? "Writing extended results... " : "Writing results... ") << endl;
    int rc;
    if (memory) {
        const int ofd = open(csv.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (ofd < 0) { cerr << "Failed to create/open file result: " << csv << endl; return; }
        rc = memory->empty() ? CUCLARK_ERR_FORMAT
                             : cuclark_classify_text_multi(c.dbs.data(), (int)c.dbs.size(), (const uint8_t*)memory->data(), memory->size(), &o,
                                                           csv_file_sink, (void*)(intptr_t)ofd, &st);
        close(ofd);
    } else {
        rc = cuclark_classify_file_multi(c.dbs.data(), (int)c.dbs.size(), objects, csv.c_str(), &o, &st);
    }
    if (rc == CUCLARK_ERR_IO) { cerr << cuclark_last_error() << endl; return; }
    if (rc == CUCLARK_ERR_FORMAT) { cerr << cuclark_last_error() << endl; exit(-1); }
    if (rc) die_lib("cuclark_classify_file");
    cerr << "Done." << endl;
    gettimeofday(&t1, nullptr);
    c.n_objects = st.n_reads;
    const double diff = (t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) / 1000000.0;
    cout << " - Assignment time: " << diff << " s. Speed: ";
    cout << (size_t)(((double)c.n_objects) / diff * 60.0) << " objects/min. (" << c.n_objects << " objects)." << endl;
    cout << " - Results stored in " << csv << endl;
}

// A whole file, mapped read-only.
struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open_path(const char* path) {
        fd = open(path, O_RDONLY);
        struct stat sb;
        if (fd < 0 || fstat(fd, &sb) != 0) return false;
        n = (size_t)sb.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~Mapped() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

// src/file.cc:205-268: FASTQ mates -> ">id\n<seq1>N<seq2>\n"; ids = first token of the header split on ' ', '/',
// '\t', '@'. The reference writes the joined reads to <file1>_ConcatenatedByCLARK.fa with getline/string code and
// classifies that file; here the two files are walked with memchr and the joined reads stay in memory.
static void merge_paired(const char* f1, const char* f2, vector<char>& out) {
    Mapped a, b;
    if (!a.open_path(f1) || !b.open_path(f2) || a.n == 0 || b.n == 0 || a.p[0] != b.p[0]) { perror("Error: the files have different format!"); exit(1); }
    if (a.p[0] != '@') { perror("Error: paired-end reads must be FASTQ files!"); exit(1); }
    out.clear();
    out.reserve(a.n / 2 + b.n / 2 + (1 << 20));
    size_t ia = 0, ib = 0;
    // [s, e) of the line at i (without the newline), i moved past it; false at end of file
    auto line = [](const Mapped& m, size_t& i, size_t& s, size_t& e) {
        if (i >= m.n) return false;
        const char* nl = (const char*)memchr(m.p + i, '\n', m.n - i);
        s = i;
        e = nl ? (size_t)(nl - m.p) : m.n;
        i = nl ? e + 1 : m.n;
        return true;
    };
    auto first_token = [](const char* p, size_t s, size_t e, size_t& ts, size_t& te) {
        auto sep = [](char c) { return c == ' ' || c == '/' || c == '\t' || c == '@'; };
        while (s < e && sep(p[s])) s++;
        ts = s;
        while (s < e && !sep(p[s])) s++;
        te = s;
    };
    size_t s1, e1, s2, e2;
    while (line(a, ia, s1, e1) && line(b, ib, s2, e2)) {
        if (e1 == s1 || e2 == s2 || a.p[s1] != '@' || b.p[s2] != '@') continue;
        size_t t1s, t1e, t2s, t2e;
        first_token(a.p, s1, e1, t1s, t1e);
        first_token(b.p, s2, e2, t2s, t2e);
        if (t1e - t1s != t2e - t2s || memcmp(a.p + t1s, b.p + t2s, t1e - t1s) != 0) { perror("Error: read id does not match between files!"); exit(1); }
        size_t q1s, q1e, q2s, q2e;
        if (!(line(a, ia, q1s, q1e) && line(b, ib, q2s, q2e))) { perror("Error: Found read without sequence"); exit(1); }
        out.push_back('>');
        out.insert(out.end(), a.p + t1s, a.p + t1e);
        out.push_back('\n');
        out.insert(out.end(), a.p + q1s, a.p + q1e);
        out.push_back('N');
        out.insert(out.end(), b.p + q2s, b.p + q2e);
        out.push_back('\n');
        size_t x, y;                                     // '+' line and qualities of both mates
        if (line(a, ia, x, y) && line(b, ib, x, y)) { if (line(a, ia, x, y) && line(b, ib, x, y)) continue; }
    }
}

static bool is_single_input(const char* objects) {
    FILE* f = fopen(objects, "r");
    string line;
    get_line(f, line);
    fclose(f);
    return !line.empty() && (line[0] == '>' || line[0] == '@' || tokens(line, " \t,", 1000).size() == 2);
}

// src/CuCLARK_hh.hh:383-427
static void run_single_end(Cli& c) {
    const bool result_exists = valid_file(c.results.c_str());
    if (!result_exists) {
        cout << "Processing file '" << c.objects << "' in " << c.batches << " batches using " << c.cpu << " CPU thread(s)." << endl;
        run_simple(c, c.objects, c.results.c_str(), false);
        return;
    }
    if (is_single_input(c.objects)) {
        cout << "Processing file'" << c.objects << "' in " << c.batches << " batches using " << c.cpu << " CPU thread(s)." << endl;
        run_simple(c, c.objects, c.results.c_str(), false);
        return;
    }
    FILE* rf = fopen(c.results.c_str(), "r");
    FILE* of = fopen(c.objects, "r");
    string ol, rl;
    cout << "Using " << c.cpu << " CPU thread(s)." << endl;
    while (get_line(of, ol) && get_line(rf, rl)) {
        cout << "> Processing file '" << ol << "' in " << c.batches << " batches." << endl;
        run_simple(c, ol.c_str(), rl.c_str(), false);
    }
    fclose(rf);
    fclose(of);
}

// src/CuCLARK_hh.hh:433-506
static void run_paired_one(Cli& c, const char* f1, const char* f2, const char* result, bool list_mode) {
    // the reference writes the joined mates to this file, classifies it and deletes it again (src/CuCLARK_hh.hh:443-451);
    // here they stay in memory, the messages keep the name
    const string merged = string(f1) + "_ConcatenatedByCLARK.fa";
    vector<char> joined;
    merge_paired(f1, f2, joined);
    if (list_mode) cout << "> Processing file: '" << merged << "' in " << c.batches << " batches." << endl;
    else cout << "Processing file: '" << merged << "' in " << c.batches << " batches using " << c.cpu << " CPU thread(s)." << endl;
    run_simple(c, merged.c_str(), result, true, &joined);
}

static void run_paired_end(Cli& c) {
    if (!valid_file(c.results.c_str()) || is_single_input(c.objects)) {
        run_paired_one(c, c.objects, c.objects2, c.results.c_str(), false);
        return;
    }
    FILE* rf = fopen(c.results.c_str(), "r");
    FILE* f1 = fopen(c.objects, "r");
    FILE* f2 = fopen(c.objects2, "r");
    string l1, l2, rl;
    cout << "Using " << c.cpu << " CPU thread(s)." << endl;
    while (get_line(f1, l1) && get_line(f2, l2) && get_line(rf, rl)) run_paired_one(c, l1.c_str(), l2.c_str(), rl.c_str(), true);
    fclose(rf);
    fclose(f1);
    fclose(f2);
}

int main(int argc, char** argv) {
    if (argc == 2) {
        const string v(argv[1]);
        if (v == "--help" || v == "--HELP") { print_usage(); return 0; }
        if (v == "--version" || v == "--VERSION") {
            cout << "Version: " << CLI_VERSION << " (Copyright 2016-2017 Robin Kobus, rkobus@students.uni-mainz.de)" << endl;
            cout << "Based on CLARK version 1.1.3 (UCR CS&E. Copyright 2013-2016 Rachid Ounit, rouni001@cs.ucr.edu) " << endl;
            cout << "B200 build (libcuclark_b200 " << cuclark_version() << ")" << endl;
            return 0;
        }
    }
    if (argc < 6) {
        cerr << "To run " << argv[0] << ", at least four  parameters are necessary:\n";
        cerr << "filename of the targets definition, directory of database, filename for objects, filename for results." << endl;
        print_usage();
        return -1;
    }
    Cli c;
    int i_targets = -1, i_objects = -1, i_objects2 = -1, i_folder = -1, i_results = -1;
    auto need = [&](int& i, const char* msg) { if (++i >= argc) { cerr << msg << endl; exit(1); } };
    for (int i = 1; i < argc; i++) {
        const string v(argv[i]);
        if (v == "-k") {
            need(i, "Please specify the k-mer length!");
            c.k = atoi(argv[i]);
            if (c.k <= 1 || c.k > MAXK) { cerr << "The k-mer length should be in [2," << MAXK << "]." << endl; exit(1); }
        } else if (v == "-t") {
            need(i, "Please specify the minimum frequency (targets)!");
            c.min_t = atoi(argv[i]);
            if (c.min_t >= 65536) { cerr << "The min k-mer frequency should be in [0,65535]." << endl; exit(1); }
        } else if (v == "-n") {
            need(i, "Please specify the number of threads!");
            c.cpu = atoi(argv[i]);
            if (c.batches < c.cpu) c.batches = c.cpu;
            if (c.cpu < 1) { cerr << "The number of threads should be higher than 0." << endl; exit(1); }
        } else if (v == "--tsk") {
            c.tsk = true;
        } else if (v == "--cache") {
            c.cache = true;
        } else if (v == "--extended") {
            c.ext = true;
        } else if (v == "-T") {
            need(i, "Please specify the targets!");
            i_targets = i;
            if (!valid_file(argv[i])) { cerr << "Failed to find/read the file of the targets definition: " << argv[i] << endl; exit(1); }
        } else if (v == "-O") {
            need(i, "Please specify the objects!");
            i_objects = i;
            if (!valid_file(argv[i])) { cerr << "Failed to find/read the filename of objects: " << argv[i] << endl; exit(1); }
        } else if (v == "-P") {
            if (i + 2 >= argc) { cerr << "Please specify the paired-end reads!" << endl; exit(1); }
            i_objects = ++i;
            i_objects2 = ++i;
            if (!valid_file(argv[i_objects])) { cerr << "Failed to find/read " << argv[i_objects] << endl; exit(1); }
            if (!valid_file(argv[i_objects2])) { cerr << "Failed to find/read " << argv[i_objects2] << endl; exit(1); }
        } else if (v == "-D") {
            need(i, "Please specify the database directory!");
            i_folder = i;
            if (!valid_file(argv[i])) { cerr << "Failed to find/read the directory:  " << argv[i] << endl; exit(1); }
        } else if (v == "-R") {
            need(i, "Please specify where to store results!");
            i_results = i;
        } else if (v == "-g") {
            need(i, "Please specify a gap value!");
            c.iter_kmers = atoi(argv[i]);
            if (c.iter_kmers < 4) { cerr << "The gap value should be >= 4." << endl; exit(1); }
        } else if (v == "-s") {
            need(i, "Please specify a sampling factor value!");
            c.sfactor = atoi(argv[i]);
            if (c.sfactor < 2 || c.sfactor > SFACTORMAX) {
                cerr << "The sampling factor value should be in the interval [2," << SFACTORMAX << "]." << endl;
                exit(1);
            }
        } else if (v == "-b") {
            need(i, "Please specify the number of batches!");
            c.batches = atoi(argv[i]);
            if (c.batches < c.cpu) { cerr << "The number of batches should be higher than the number of threads." << endl; exit(1); }
        } else if (v == "-d") {
            need(i, "Please specify the number of devices to use!");
            c.devices = atoi(argv[i]);
            if (c.devices < 1) { cerr << "The number of devices should be higher than 0." << endl; exit(1); }
        } else {
            cerr << "Failed to recognize option: " << v << endl;
            exit(1);
        }
    }
    if (HTSIZE == CUCLARK_HTSIZE_LIGHT) {      // src/main.cc:241-253
        c.light = true;
        if (c.iter_kmers == 0) c.iter_kmers = 4;
        c.k = 27;
        c.sfactor = 1;
    } else {
        c.iter_kmers = 0;
    }
    if (i_targets < 0 || i_folder < 0 || i_objects < 0 || i_results < 0) {
        cerr << "Failed to run " << argv[0] << ": at least four  parameters are necessary";
        cerr << ": file of targets, directory of database, file of objects, file for results." << endl;
        print_usage();
        exit(1);
    }
    c.targets = argv[i_targets];
    c.folder = argv[i_folder];
    if (c.folder.back() != '/') c.folder.push_back('/');
    c.objects = argv[i_objects];
    c.objects2 = i_objects2 > 0 ? argv[i_objects2] : nullptr;
    c.results = argv[i_results];

    read_targets(c);
    banner(c);
    load_database(c);
    if (c.objects2) run_paired_end(c);
    else run_single_end(c);
    for (cuclark_db* db : c.dbs) cuclark_destroy(db);
    return 0;
}
