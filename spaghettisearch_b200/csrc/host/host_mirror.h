// Host-side mirror of the reference's Go API for the two hot paths, in C++ because no
// Go toolchain exists in this image (the cgo version of the same code is INTEGRATION.md).
//
// Same names, argument meaning and error behaviour as the reference:
//   ranking.UpdateTopicSensitivePagerank   ranking/pagerank.go:14
//   ranking.UpdateTermWeights              ranking/term_weighting.go:10
//   retrieval.Retrieve                     retrieval/main_retrieve.go:15
// over in-memory snapshots of the Badger tables in the reference's own value
// encodings (database/noschema_schema.go:125-260, database/database.go:85-122):
//   forw[2]  docHash  -> ["childHash", ...] | null
//   forw[3]  docHash  -> {category: rank}
//   forw[4]  docHash  -> {"title": norm, "body": norm}
//   forw[5]  category -> {"numPages": n, "wordCount": w}
//   inv[0|1] wordHash -> {docHash: [weight, pos...]}
// Failures throw std::runtime_error -- the reference panics (pagerank.go:20,29,49,...).
// Everything between export and write-back goes through the C ABI of include/spaghetti.h.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "spaghetti.h"

namespace db {

// One Badger table: ordered key -> JSON value bytes (database/database.go:42-75).
class Table {
 public:
  using Rows = std::map<std::string, std::string>;
  // Get: value or nullptr when the key is absent (badger.ErrKeyNotFound)
  const std::string* Get(const std::string& key) const;
  void Set(const std::string& key, std::string value);
  const Rows& Iterate() const { return rows_; }  // the reference's order is random; here ascending key
  size_t size() const { return rows_.size(); }
  void DropTable() { rows_.clear(); }
  // JSON-lines snapshot: {"k": "<key>", "v": <value>} per row
  void LoadJsonl(const std::string& path);
  void SaveJsonl(const std::string& path) const;

 private:
  Rows rows_;
};

// database.DB_init -> inv[0..2], forw[0..5] (database/database.go:101-122)
struct DB {
  Table inv[3];
  Table forw[6];
};

// value codecs (encoding/json shapes used by the hot paths)
std::vector<std::string> ParseStringArray(const std::string& json);                          // []string | null
std::vector<std::pair<std::string, double>> ParseNumberMap(const std::string& json);        // map[string]float64
std::vector<std::pair<std::string, std::vector<float>>> ParsePostings(const std::string& json);  // map[string][]float32
std::string FormatNumberMap(const std::vector<std::pair<std::string, double>>& m);
std::string FormatPostings(const std::vector<std::pair<std::string, std::vector<float>>>& m);

}  // namespace db

namespace ranking {

// CSR export of forw[2] on dense ids (id = rank of the hex key among parents U children).
struct GraphExport {
  std::vector<std::string> keys;   // dense id -> docHash
  std::vector<uint64_t> row_ptr;   // [N+1]
  std::vector<uint32_t> col_idx;   // [E]
};
GraphExport ExportGraph(const db::Table& forw2);  // pagerank.go:18-44

void UpdateTopicSensitivePagerank(ss_engine* e, double dampingFactor, double convergenceCriterion,
                                  db::Table forward[6]);
void UpdateTermWeights(ss_engine* e, db::Table* inv, db::Table forw[6], const std::string& info);

}  // namespace ranking

namespace retrieval {

struct Rank_combined {  // retrieval/util.go:25-36, the fields the hot path fills
  std::string DocHash;
  double PageRank = 0;
  double FinalRank = 0;
};

// The weighted tables on the device, loaded once (a server keeps this alive).
class Index {
 public:
  Index(ss_engine* e, db::Table forw[6], db::Table inv[3]);
  // tokens are md5-hex hashes of the laundered words (main_retrieve.go:28-36); duplicates kept
  std::vector<Rank_combined> Retrieve(const std::vector<std::string>& queryTokenised,
                                      const std::vector<std::string>& phraseTokenised, uint32_t k = 50) const;
  // a whole batch in one ss_score_batch call; results[i] belongs to queries[i]
  struct Query {
    std::vector<std::string> queryTokenised, phraseTokenised;
  };
  std::vector<std::vector<Rank_combined>> RetrieveBatch(const std::vector<Query>& queries, uint32_t k = 50) const;

 private:
  ss_engine* e_;
  std::vector<std::string> doc_keys_;
  std::map<std::string, uint32_t> term_id_;
};

// Query front-end batching: retrieval.Retrieve is called concurrently, one
// goroutine per HTTP request (cmd/server/server.go:32-52), but the scoring kernel wants batches.
// Submit() may be called from any number of threads; a worker collects what arrived within
// `window_us` (or `max_batch` queries), scores them in one ss_score_batch call and hands each
// caller its own result.  Same results as Index::Retrieve, query by query.
class BatchingRetriever {
 public:
  BatchingRetriever(const Index& index, uint32_t k = 50, uint32_t max_batch = 4096, uint32_t window_us = 200);
  ~BatchingRetriever();
  std::vector<Rank_combined> Submit(const std::vector<std::string>& queryTokenised,
                                    const std::vector<std::string>& phraseTokenised);
  uint64_t batches() const;   // ss_score_batch calls so far
  uint64_t queries() const;   // queries served so far

 private:
  struct Impl;
  Impl* impl_;
};

std::vector<Rank_combined> Retrieve(ss_engine* e, const std::vector<std::string>& queryTokenised,
                                    const std::vector<std::string>& phraseTokenised, db::Table forw[6],
                                    db::Table inv[3]);

}  // namespace retrieval
