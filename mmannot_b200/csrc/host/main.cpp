// mmannot_b200 -- drop-in command line for mmannot's read-annotation path on a B200.
// Flags, defaults, messages and exit codes follow main() of the reference
// (mmannot.cpp:1903-2149); the annotation itself runs on the GPU through the C ABI
// (include/mmannot_b200.h).  There is no CPU fallback.
#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>
#include <string>
#include <vector>

#include "annotation.hpp"
#include "config.hpp"
#include "counter.hpp"
#include "mmannot_b200.h"

using namespace mmb;

namespace {

const char VERSION[] = "1.1";

void printUsage() {
  std::cerr << "Usage: mmannot [options]\n"
               "\tCompulsory options:\n"
               "\t\t-a file: annotation file in GTF format\n"
               "\t\t-r file1 [file2 ...]: reads in BAM/SAM format\n"
               "\tMain options:\n"
               "\t\t-o output: output file (default: stdout)\n"
               "\t\t-c config_file: configuration file (default: config.txt)\n"
               "\t\t-n name1 name2...: short name for each of the reads files\n"
               "\t\t-s strand: string (U, F, R, FR, RF, FF, defaut: F) (use several strand types if the library strategies differ)\n"
               "\t\t-f format (SAM or BAM): format of the read files (default: guess from file extension)\n"
               "\t\t-l integer: overlap type (<0: read is included, <1: % overlap, otherwise: # nt, default: -1)\n"
               "\t\t-d integer: upstream region size (default: 1000)\n"
               "\t\t-D integer: downstream region size (default: 1000)\n"
               "\t\t-y string: quantification strategy, valid values are: default, unique, random, ratio (default: default)\n"
               "\t\t-e integer: attribute a read to a feature if at least N% of the hits map to the feature (default: 100%)\n"
               "\tOutput options:\n"
               "\t\t-p: print progress\n"
               "\t\t-m file: print mapping statistics for each read (slow, only work with 1 input file)\n"
               "\t\t-M file: print mapping statistics for each interval (slow, only work with 1 input file)\n"
               "\t\t-t integer: # threads (default: 1)\n"
               "\t\t-g integer: CUDA device (default: 0)\n"
               "\t\t-G integer: # GPUs every input is spread over, by read name (default: 1)\n"
               "\t\t-h: this help"
            << std::endl;
}

int fail(const std::string &msg) {
  std::cerr << msg << std::endl;
  return EXIT_FAILURE;
}

}  // namespace

int main(int argc, char **argv) {
  RunOptions opt;
  AnnotationOptions annOpt;
  std::string gtfFileName, outputFileName, configFileName = "config.txt", readStatsFile, intervalStatsFile;
  std::vector<std::string> readsFileNames, names;
  int device = 0, nThreads = 1, gpusPerInput = 1;
  if (argc == 1) {
    printUsage();
    return EXIT_SUCCESS;
  }
  auto value = [&](int &i) -> std::string {
    if (i + 1 >= argc) { std::cerr << "Missing value after '" << argv[i] << "'.\nExiting." << std::endl; printUsage(); std::exit(EXIT_FAILURE); }
    return std::string(argv[++i]);
  };
  for (int i = 1; i < argc; i++) {
    std::string s(argv[i]);
    if (s.empty()) continue;
    if (s == "-a") gtfFileName = value(i);
    else if (s == "-r" || s == "-n") {
      std::vector<std::string> &dst = (s == "-r") ? readsFileNames : names;
      for (++i; i < argc; ++i) {
        std::string t(argv[i]);
        if (!t.empty() && t[0] == '-') { --i; break; }
        dst.push_back(t);
      }
    }
    else if (s == "-c") configFileName = value(i);
    else if (s == "-o") outputFileName = value(i);
    else if (s == "-l") opt.overlap = std::stof(value(i));
    else if (s == "-s") {
      for (++i; i < argc; ++i) {
        std::string t(argv[i]);
        if (t == "U") opt.strandedness = Strandedness::U;
        else if (t == "F") opt.strandedness = Strandedness::F;
        else if (t == "R") opt.strandedness = Strandedness::R;
        else if (t.empty() || t[0] == '-') { --i; break; }
        else {
          std::cerr << "Do not understand strandedness " << t << "\nExiting." << std::endl;
          printUsage();
          return EXIT_FAILURE;
        }
      }
    }
    else if (s == "-p") opt.progress = true;
    else if (s == "-t") nThreads = std::max(1, std::stoi(value(i)));  // workers, one input file at a time each, spread over the GPUs
    else if (s == "-g") device = std::stoi(value(i));
    else if (s == "-G") gpusPerInput = std::max(1, std::stoi(value(i)));
    else if (s == "-m") { readStatsFile = value(i); opt.readStats = true; }
    else if (s == "-M") { intervalStatsFile = value(i); opt.intervalStats = true; }
    else if (s == "-f") {
      for (++i; i < argc; ++i) {
        std::string t = lowered(argv[i]);
        if (t == "sam") opt.format = ReadsFormat::SAM;
        else if (t == "bam") opt.format = ReadsFormat::BAM;
        else if (t.empty() || t[0] == '-') { --i; break; }
        else {
          std::cerr << "Do not understand reads format " << t << "\nExiting." << std::endl;
          printUsage();
          return EXIT_FAILURE;
        }
      }
    }
    else if (s == "-e") opt.rescueThreshold = static_cast<float>(std::stof(value(i)) / 100.0);  // mm:2024
    else if (s == "-d") annOpt.upstreamSize = std::stoul(value(i));
    else if (s == "-D") annOpt.downstreamSize = std::stoul(value(i));
    else if (s == "-y") {
      std::string t = lowered(value(i));
      if (t == "default") opt.strategy = MMA_STRATEGY_DEFAULT;
      else if (t == "unique") opt.strategy = MMA_STRATEGY_UNIQUE;
      else if (t == "random") opt.strategy = MMA_STRATEGY_RANDOM;
      else if (t == "ratio") opt.strategy = MMA_STRATEGY_RATIO;
      else {
        std::cerr << "Do not understand strategy " << t << "\nExiting." << std::endl;
        printUsage();
        return EXIT_FAILURE;
      }
    }
    else if (s == "-v") { std::cerr << "mmannot v" << VERSION << std::endl; return EXIT_SUCCESS; }
    else if (s == "-h") { printUsage(); return EXIT_SUCCESS; }
    else {
      std::cerr << "Error: wrong parameter '" << s << "'.\nExiting." << std::endl;
      printUsage();
      return EXIT_FAILURE;
    }
  }
  if (gtfFileName.empty()) { std::cerr << "Missing input GTF file.\nExiting." << std::endl; printUsage(); return EXIT_FAILURE; }
  if (readsFileNames.empty()) { std::cerr << "Missing input BAM file.\nExiting." << std::endl; printUsage(); return EXIT_FAILURE; }
  const uint32_t nInputs = static_cast<uint32_t>(readsFileNames.size());
  if (names.empty()) {
    for (const std::string &fileName : readsFileNames) {  // basename without its last extension, mm:2072-2081
      std::string n = fileName;
      size_t p = n.find_last_of("/");
      if (p != std::string::npos) n = n.substr(p + 1);
      p = n.find_last_of(".");
      if (p != std::string::npos) n = n.substr(0, p);
      names.push_back(n);
    }
  } else if (names.size() != nInputs) {
    std::cerr << "Number of names is not equal to number of file names.\nExiting." << std::endl;
    printUsage();
    return EXIT_FAILURE;
  }
  if ((opt.readStats || opt.intervalStats) && nInputs != 1) {
    std::cerr << "Only one reads file when providing reads or interval statistics.\nExiting." << std::endl;
    printUsage();
    return EXIT_FAILURE;
  }
  // the CUDA context comes up on a side thread while the configuration and the annotation are parsed
  std::thread warm([device, nThreads, gpusPerInput]() {
    const int nDev = std::max(1, mma_device_count());
    for (int w = 0; w < std::min(nThreads * gpusPerInput, nDev); ++w) mma_warmup((device + w) % nDev);
  });
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joinWarm{warm};
  std::string err, warnings;
  Config config;
  if (!config.parse(configFileName, err)) return fail(err);
  std::cerr << config.orderEcho();
  std::cerr << "Reading GTF file" << std::endl;
  FeatureTable features;
  bool ok = buildFeatureTable(gtfFileName, config, annOpt, features, err, warnings);
  std::cerr << warnings;
  if (!ok && features.nLines == 0 && features.nGenes == 0 && err.find("not been parsed properly") == std::string::npos) return fail(err);
  std::cerr << "\t" << withThousands(features.nLines) << " lines read, done.  " << withThousands(features.nGenes) << " genes found." << std::endl;
  if (!ok) return fail(err);
  std::cerr << "\t" << withThousands(features.size()) << " intervals found." << std::endl;

  std::ofstream of;
  if (!outputFileName.empty()) of.open(outputFileName.c_str());
  std::ostream &outputFile = outputFileName.empty() ? std::cout : of;

  std::vector<uint16_t> elemLine;
  std::vector<uint8_t> elemStrand, elemVic;
  config.deviceTables(elemLine, elemStrand, elemVic);
  mma_params params = mma_params();
  params.device = device;
  params.strategy = opt.strategy;
  params.overlap = opt.overlap;
  params.rescue_threshold = opt.rescueThreshold;
  params.read_stats = opt.readStats;
  params.interval_stats = opt.intervalStats;
  params.n_elements = static_cast<uint32_t>(elemLine.size());
  params.elem_line = elemLine.data();
  params.elem_strand = elemStrand.data();
  params.elem_vicinity = elemVic.data();
  params.n_samples = nInputs;
  params.max_batch_hits = opt.batchHits;
  params.table_log2 = 20;  // 2^20 combination slots per input file (16 MB): heavy multi-mapping can produce many distinct sets
  params.rand_seed = 1;
  mma_features f;
  f.n = static_cast<uint32_t>(features.size());
  f.n_chr = static_cast<uint32_t>(features.chromosomes.size());
  f.chr = features.chr.data(); f.start = features.start.data(); f.end = features.end.data();
  f.type = features.type.data(); f.strand = features.strand.data();

  // -t n: n workers take the input files in turn (the reference's pool, mm:2117-2141: one thread per file); worker w drives its
  // own context on GPU (device + w) mod #GPUs, with the feature index replicated.  Reports and table columns keep file order.
  const int nDevices = std::max(1, mma_device_count());
  // -G n: every input is spread over n GPUs by read name (all the records of a name on one GPU), tables summed on the devices
  const uint32_t nShards = (opt.readStats || opt.intervalStats) ? 1u : static_cast<uint32_t>(std::min(gpusPerInput, nDevices));
  const uint32_t nWorkers = (opt.readStats || opt.intervalStats)
                                ? 1u
                                : std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(static_cast<uint32_t>(nThreads), nInputs),
                                                                         nShards > 1 ? static_cast<uint32_t>(nDevices) / nShards : 0xFFFFFFFFu));
  struct FileResult { bool ok = false; std::string err, log; std::map<uint64_t, double> counts; };
  std::vector<FileResult> results(nInputs);
  std::ofstream readStatsStream, intervalStatsStream;
  if (opt.readStats) readStatsStream.open(readStatsFile.c_str());            // mm:2000
  if (opt.intervalStats) intervalStatsStream.open(intervalStatsFile.c_str());  // mm:2005
  StatsWriters writers(config, features, opt.strategy, opt.rescueThreshold, opt.readStats ? &readStatsStream : nullptr, opt.intervalStats);
  std::string fatal;
  std::mutex fatalMutex;
  auto worker = [&](uint32_t w) {
    std::vector<mma_ctx *> ctxs;
    auto destroyAll = [&ctxs]() { for (mma_ctx *c : ctxs) mma_destroy(c); ctxs.clear(); };
    for (uint32_t g = 0; g < nShards; ++g) {
      mma_params p = params;
      p.device = (device + static_cast<int>(w * nShards + g)) % nDevices;
      mma_ctx *ctx = nullptr;
      if (mma_create(&ctx, &p) != MMA_OK) {
        std::lock_guard<std::mutex> lk(fatalMutex);
        fatal = std::string("Error: ") + mma_last_error(nullptr);
        destroyAll();
        return;
      }
      ctxs.push_back(ctx);
      if (mma_load_features(ctx, &f) != MMA_OK) {
        std::lock_guard<std::mutex> lk(fatalMutex);
        fatal = std::string("Error: ") + mma_last_error(ctx);
        destroyAll();
        return;
      }
    }
    {
      Counter counter(ctxs, features, config, opt);
      if (opt.readStats || opt.intervalStats) counter.setStatsWriters(&writers);
      for (uint32_t i = w; i < nInputs; i += nWorkers) {
        FileResult &r = results[i];
        std::ostringstream log;
        std::ostream &lg = (nWorkers == 1) ? static_cast<std::ostream &>(std::cerr) : static_cast<std::ostream &>(log);
        const auto tRead0 = std::chrono::steady_clock::now();
        r.ok = counter.read(readsFileNames[i], i, r.err, lg);
        if (std::getenv("MMANNOT_B200_TIMING"))  // (bench.py: time of Counter::read alone, without annotation load and CUDA start-up)
          lg << "[timing] read_ms=" << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tRead0).count() << " file=" << readsFileNames[i] << std::endl;
        if (r.ok) {
          counter.dump(lg);
          if (opt.intervalStats) writers.dumpIntervals(intervalStatsStream);
          r.counts = counter.getCounts();
        }
        r.log = log.str();
        if (!r.ok) break;
      }
    }
    destroyAll();
  };
  if (nWorkers == 1) worker(0);
  else {
    std::vector<std::thread> pool;
    for (uint32_t w = 0; w < nWorkers; ++w) pool.emplace_back(worker, w);
    for (std::thread &t : pool) t.join();
  }
  if (!fatal.empty()) return fail(fatal);
  int rc = 0;
  {
    TableCount table(config, nInputs);
    for (uint32_t i = 0; i < nInputs && rc == 0; i++) {
      std::cerr << results[i].log;
      if (!results[i].ok) { std::cerr << results[i].err << std::endl; rc = EXIT_FAILURE; break; }
      table.addCounts(results[i].counts);
    }
    if (rc == 0) table.dump(outputFile, names);
  }
  if (rc == 0) std::cerr << "Successfully done." << std::endl;
  return rc;
}
