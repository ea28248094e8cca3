// See host_mirror.h.  Export (tables -> dense CSR/CSC), C-ABI calls, write-back.
#include "host_mirror.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <future>
#include <mutex>
#include <thread>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error(msg); }

void must(int rc, const char* what) {  // the Go shim panics with ss_last_error(); so do we
  if (rc < 0) fail(std::string(what) + ": " + ss_last_error());
}

// ---- a small JSON reader for the value shapes the tables use -----------------------
struct Cursor {
  const char* p;
  const char* end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  void expect(char c) {
    if (!eat(c)) fail(std::string("json: expected '") + c + "'");
  }
  bool literal(const char* lit) {
    ws();
    size_t n = strlen(lit);
    if ((size_t)(end - p) >= n && !memcmp(p, lit, n)) {
      p += n;
      return true;
    }
    return false;
  }
  std::string str() {
    expect('"');
    std::string out;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) fail("json: bad escape");
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (end - p < 5) fail("json: bad \\u");
            unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
            p += 4;
            if (cp < 0x80) {
              out += (char)cp;
            } else if (cp < 0x800) {
              out += (char)(0xC0 | (cp >> 6));
              out += (char)(0x80 | (cp & 0x3F));
            } else {
              out += (char)(0xE0 | (cp >> 12));
              out += (char)(0x80 | ((cp >> 6) & 0x3F));
              out += (char)(0x80 | (cp & 0x3F));
            }
            break;
          }
          default: out += *p;  // \" \\ \/
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= end) fail("json: unterminated string");
    ++p;
    return out;
  }
  // numeric token as text (so that float32 and float64 targets round once, like Go's ParseFloat)
  std::string number() {
    ws();
    const char* b = p;
    while (p < end && (isdigit((unsigned char)*p) || *p == '-' || *p == '+' || *p == '.' || *p == 'e' || *p == 'E'))
      ++p;
    if (p == b) fail("json: expected a number");
    return std::string(b, p);
  }
  // skip any value, returning its raw text
  std::string raw() {
    ws();
    const char* b = p;
    int depth = 0;
    bool in_str = false;
    for (; p < end; ++p) {
      if (in_str) {
        if (*p == '\\') ++p;
        else if (*p == '"') in_str = false;
        continue;
      }
      if (*p == '"') in_str = true;
      else if (*p == '{' || *p == '[') ++depth;
      else if (*p == '}' || *p == ']') {
        if (depth == 0) break;
        --depth;
      } else if (*p == ',' && depth == 0) break;
    }
    const char* e = p;
    while (e > b && (e[-1] == ' ' || e[-1] == '\n' || e[-1] == '\r' || e[-1] == '\t')) --e;
    return std::string(b, e);
  }
};

std::string quote(const std::string& s) {
  std::string o = "\"";
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') {
      o += '\\';
      o += (char)c;
    } else if (c < 0x20) {
      char buf[8];
      snprintf(buf, sizeof(buf), "\\u%04x", c);
      o += buf;
    } else {
      o += (char)c;
    }
  }
  return o + "\"";
}
std::string fmt_double(double v) {  // round-trips; Go writes the shortest form of the same value
  if (std::isnan(v) || std::isinf(v)) fail("json: unsupported value (NaN/Inf), as in encoding/json");
  char buf[40];
  snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}
std::string fmt_float(float v) {
  if (std::isnan(v) || std::isinf(v)) fail("json: unsupported value (NaN/Inf), as in encoding/json");
  char buf[32];
  snprintf(buf, sizeof(buf), "%.9g", (double)v);
  return buf;
}

}  // namespace

namespace db {

const std::string* Table::Get(const std::string& key) const {
  auto it = rows_.find(key);
  return it == rows_.end() ? nullptr : &it->second;
}
void Table::Set(const std::string& key, std::string value) { rows_[key] = std::move(value); }

void Table::LoadJsonl(const std::string& path) {
  std::ifstream in(path);
  if (!in) fail("cannot open " + path);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    Cursor c{line.data(), line.data() + line.size()};
    c.expect('{');
    std::string key, val;
    bool have_k = false, have_v = false;
    do {
      std::string name = c.str();
      c.expect(':');
      if (name == "k") {
        key = c.str();
        have_k = true;
      } else if (name == "v") {
        val = c.raw();
        have_v = true;
      } else {
        c.raw();
      }
    } while (c.eat(','));
    c.expect('}');
    if (!have_k || !have_v) fail("jsonl row without k/v in " + path);
    rows_[key] = val;
  }
}
void Table::SaveJsonl(const std::string& path) const {
  std::ofstream out(path);
  if (!out) fail("cannot write " + path);
  for (auto& kv : rows_) out << "{\"k\": " << quote(kv.first) << ", \"v\": " << kv.second << "}\n";
}

std::vector<std::string> ParseStringArray(const std::string& json) {
  Cursor c{json.data(), json.data() + json.size()};
  std::vector<std::string> out;
  if (c.literal("null")) return out;
  c.expect('[');
  if (c.eat(']')) return out;
  do out.push_back(c.str());
  while (c.eat(','));
  c.expect(']');
  return out;
}
std::vector<std::pair<std::string, double>> ParseNumberMap(const std::string& json) {
  Cursor c{json.data(), json.data() + json.size()};
  std::vector<std::pair<std::string, double>> out;
  if (c.literal("null")) return out;
  c.expect('{');
  if (c.eat('}')) return out;
  do {
    std::string k = c.str();
    c.expect(':');
    out.emplace_back(k, strtod(c.number().c_str(), nullptr));
  } while (c.eat(','));
  c.expect('}');
  return out;
}
std::vector<std::pair<std::string, std::vector<float>>> ParsePostings(const std::string& json) {
  Cursor c{json.data(), json.data() + json.size()};
  std::vector<std::pair<std::string, std::vector<float>>> out;
  if (c.literal("null")) return out;
  c.expect('{');
  if (c.eat('}')) return out;
  do {
    std::string k = c.str();
    c.expect(':');
    std::vector<float> v;
    if (!c.literal("null")) {
      c.expect('[');
      if (!c.eat(']')) {
        do v.push_back(strtof(c.number().c_str(), nullptr));
        while (c.eat(','));
        c.expect(']');
      }
    }
    out.emplace_back(std::move(k), std::move(v));
  } while (c.eat(','));
  c.expect('}');
  return out;
}
std::string FormatNumberMap(const std::vector<std::pair<std::string, double>>& m) {
  std::string o = "{";
  for (size_t i = 0; i < m.size(); ++i) o += (i ? "," : "") + quote(m[i].first) + ":" + fmt_double(m[i].second);
  return o + "}";
}
std::string FormatPostings(const std::vector<std::pair<std::string, std::vector<float>>>& m) {
  std::string o = "{";
  for (size_t i = 0; i < m.size(); ++i) {
    o += (i ? "," : "") + quote(m[i].first) + ":[";
    for (size_t j = 0; j < m[i].second.size(); ++j) o += (j ? "," : "") + fmt_float(m[i].second[j]);
    o += "]";
  }
  return o + "}";
}

}  // namespace db

namespace {

// dense ids = rank in ascending key order
std::map<std::string, uint32_t> dense_ids(const std::vector<std::string>& sorted_keys) {
  std::map<std::string, uint32_t> m;
  for (size_t i = 0; i < sorted_keys.size(); ++i) m.emplace_hint(m.end(), sorted_keys[i], (uint32_t)i);
  return m;
}

struct PostingExport {  // one inverted table, term major, docs ascending inside a term
  std::vector<std::string> terms;
  std::vector<uint64_t> term_ptr{0};
  std::vector<uint32_t> doc_ids;
  std::vector<float> w;
  std::vector<uint64_t> pos_ptr{0};
  std::vector<float> pos;
};

PostingExport export_postings(const db::Table& inv, const std::vector<std::string>& term_keys,
                              const std::map<std::string, uint32_t>& doc_id) {
  PostingExport x;
  x.terms = term_keys;
  for (auto& term : term_keys) {
    const std::string* val = inv.Get(term);
    if (val) {
      auto row = db::ParsePostings(*val);
      std::vector<std::pair<uint32_t, const std::vector<float>*>> ents;
      ents.reserve(row.size());
      for (auto& kv : row) {
        auto it = doc_id.find(kv.first);
        if (it == doc_id.end()) fail("posting for unknown doc " + kv.first);
        if (kv.second.empty()) fail("posting without a weight for doc " + kv.first);  // listPos[0] would panic
        ents.emplace_back(it->second, &kv.second);
      }
      std::sort(ents.begin(), ents.end(), [](auto& a, auto& b) { return a.first < b.first; });
      for (auto& en : ents) {
        x.doc_ids.push_back(en.first);
        x.w.push_back((*en.second)[0]);
        x.pos.insert(x.pos.end(), en.second->begin() + 1, en.second->end());
        x.pos_ptr.push_back(x.pos.size());
      }
    }
    x.term_ptr.push_back(x.doc_ids.size());
  }
  return x;
}

template <class T>
const T* ptr_or_null(const std::vector<T>& v) {
  return v.empty() ? nullptr : v.data();
}

}  // namespace

namespace ranking {

GraphExport ExportGraph(const db::Table& forw2) {
  GraphExport g;
  std::set<std::string> all;  // webNodesAll, pagerank.go:24-39
  std::map<std::string, std::vector<std::string>> children;
  for (auto& kv : forw2.Iterate()) {
    auto kids = db::ParseStringArray(kv.second);
    for (auto& c : kids) all.insert(c);
    all.insert(kv.first);
    children.emplace(kv.first, std::move(kids));
  }
  g.keys.assign(all.begin(), all.end());
  auto id = dense_ids(g.keys);
  g.row_ptr.assign(1, 0);
  for (auto& k : g.keys) {
    auto it = children.find(k);
    if (it != children.end())
      for (auto& c : it->second) g.col_idx.push_back(id.at(c));  // list entries, duplicates included (:140-142)
    g.row_ptr.push_back(g.col_idx.size());
  }
  return g;
}

void UpdateTopicSensitivePagerank(ss_engine* e, double dampingFactor, double convergenceCriterion,
                                  db::Table forward[6]) {
  GraphExport g = ExportGraph(forward[2]);
  // categories, pagerank.go:47-61
  std::vector<std::string> cats;
  std::vector<int64_t> num_pages;
  for (auto& kv : forward[5].Iterate()) {
    double np = 0;
    for (auto& f : db::ParseNumberMap(kv.second))
      if (f.first == "numPages") np = f.second;
    cats.push_back(kv.first);
    num_pages.push_back((int64_t)np);  // int(val["numPages"]), :61
  }
  const size_t N = g.keys.size(), T = cats.size();
  must(ss_graph_load_csr(e, N, g.col_idx.size(), g.row_ptr.data(), ptr_or_null(g.col_idx)), "ss_graph_load_csr");
  std::vector<double> rank(N * T);
  for (size_t lo = 0; lo < T; lo += 16) {  // topics are independent runs (pagerank.go:54-63)
    const size_t hi = std::min(T, lo + 16), w = hi - lo;
    std::vector<double> slab(N * w);
    must(ss_pagerank(e, dampingFactor, convergenceCriterion, (uint32_t)w, num_pages.data() + lo, 0,
                     N ? slab.data() : nullptr, nullptr),
         "ss_pagerank");
    for (size_t v = 0; v < N; ++v) std::copy(slab.begin() + v * w, slab.begin() + (v + 1) * w, rank.begin() + v * T + lo);
  }
  // forw[3][node] = {category: rank} for every node (:66-82); {} when forw[5] is empty
  for (size_t v = 0; v < N; ++v) {
    std::vector<std::pair<std::string, double>> pr;
    for (size_t t = 0; t < T; ++t) pr.emplace_back(cats[t], rank[v * T + t]);
    forward[3].Set(g.keys[v], db::FormatNumberMap(pr));
  }
}

void UpdateTermWeights(ss_engine* e, db::Table* inv, db::Table forw[6], const std::string& info) {
  if (info != "title" && info != "body") fail("UpdateTermWeights: info must be \"title\" or \"body\"");
  const double totalDocs = (double)forw[3].size();  // term_weighting.go:12-17
  // doc id space: every doc hash seen in forw[3] or in this table
  std::set<std::string> docs;
  for (auto& kv : forw[3].Iterate()) docs.insert(kv.first);
  std::vector<std::string> terms;
  for (auto& kv : inv->Iterate()) {
    terms.push_back(kv.first);
    for (auto& p : db::ParsePostings(kv.second)) docs.insert(p.first);
  }
  std::vector<std::string> doc_keys(docs.begin(), docs.end());
  auto doc_id = dense_ids(doc_keys);
  PostingExport x = export_postings(*inv, terms, doc_id);
  const int table = info == "title" ? SS_TITLE : SS_BODY;
  const size_t P = x.doc_ids.size(), D = doc_keys.size();
  must(ss_index_clear(e), "ss_index_clear");
  must(ss_index_load(e, table, terms.size(), D, x.term_ptr.data(), ptr_or_null(x.doc_ids), ptr_or_null(x.w),
                     x.pos_ptr.data(), ptr_or_null(x.pos)),
       "ss_index_load");
  std::vector<float> w(P);
  std::vector<double> mag(D);
  must(ss_term_weights(e, table, totalDocs, nullptr, P ? w.data() : nullptr, D ? mag.data() : nullptr),
       "ss_term_weights");
  // write the rows back with listPos[0] = tf-idf (:42-47)
  std::vector<char> in_table(D, 0);  // pageMagnitude has an entry only for docs of this table
  for (size_t t = 0; t < terms.size(); ++t) {
    std::vector<std::pair<std::string, std::vector<float>>> row;
    for (uint64_t p = x.term_ptr[t]; p < x.term_ptr[t + 1]; ++p) {
      std::vector<float> list{w[p]};
      list.insert(list.end(), x.pos.begin() + x.pos_ptr[p], x.pos.begin() + x.pos_ptr[p + 1]);
      row.emplace_back(doc_keys[x.doc_ids[p]], std::move(list));
      in_table[x.doc_ids[p]] = 1;
    }
    inv->Set(terms[t], db::FormatPostings(row));
  }
  // saveMagnitude (:59-123): existing rows get `info` merged (0 when the doc is absent from this
  // table, sqrt(0)); docs of this table without a row get a new one
  db::Table& f4 = forw[4];
  auto merged = [&](const std::string& key, double m) {
    std::vector<std::pair<std::string, double>> row;
    if (const std::string* old = f4.Get(key)) row = db::ParseNumberMap(*old);
    bool found = false;
    for (auto& kv : row)
      if (kv.first == info) {
        kv.second = m;
        found = true;
      }
    if (!found) row.emplace_back(info, m);
    return db::FormatNumberMap(row);
  };
  std::vector<std::string> existing;
  for (auto& kv : f4.Iterate()) existing.push_back(kv.first);
  for (auto& key : existing) {
    auto it = doc_id.find(key);
    f4.Set(key, merged(key, it != doc_id.end() && in_table[it->second] ? mag[it->second] : 0.0));
  }
  for (size_t d = 0; d < D; ++d)
    if (in_table[d] && !f4.Get(doc_keys[d])) f4.Set(doc_keys[d], merged(doc_keys[d], mag[d]));
}

}  // namespace ranking

namespace retrieval {

Index::Index(ss_engine* e, db::Table forw[6], db::Table inv[3]) : e_(e) {
  std::set<std::string> docs, terms;
  for (auto& kv : forw[3].Iterate()) docs.insert(kv.first);
  for (auto& kv : forw[4].Iterate()) docs.insert(kv.first);
  for (int t = 0; t < 2; ++t)
    for (auto& kv : inv[t].Iterate()) {
      terms.insert(kv.first);
      for (auto& p : db::ParsePostings(kv.second)) docs.insert(p.first);
    }
  doc_keys_.assign(docs.begin(), docs.end());
  auto doc_id = dense_ids(doc_keys_);
  std::vector<std::string> term_keys(terms.begin(), terms.end());
  term_id_ = dense_ids(term_keys);
  const size_t D = doc_keys_.size();
  must(ss_index_clear(e_), "ss_index_clear");
  for (int t = 0; t < 2; ++t) {  // inv[0] = title, inv[1] = body; weights are already tf-idf
    PostingExport x = export_postings(inv[t], term_keys, doc_id);
    must(ss_index_load(e_, t, term_keys.size(), D, x.term_ptr.data(), ptr_or_null(x.doc_ids), ptr_or_null(x.w),
                       x.pos_ptr.data(), ptr_or_null(x.pos)),
         "ss_index_load");
    std::vector<double> mag(D, 0.0);  // forw[4][doc]["title"|"body"], absent key reads as 0
    for (auto& kv : forw[4].Iterate())
      for (auto& f : db::ParseNumberMap(kv.second))
        if (f.first == (t == 0 ? "title" : "body")) mag[doc_id.at(kv.first)] = f.second;
    must(ss_set_doc_norms(e_, t, D, D ? mag.data() : nullptr), "ss_set_doc_norms");
  }
}

std::vector<std::vector<Rank_combined>> Index::RetrieveBatch(const std::vector<Query>& queries, uint32_t k) const {
  auto ids = [&](const std::vector<std::string>& toks, std::vector<uint32_t>& out) {
    for (auto& t : toks) {
      auto it = term_id_.find(t);
      out.push_back(it == term_id_.end() ? 0xFFFFFFFFu : it->second);  // ErrKeyNotFound => empty row
    }
  };
  const size_t Q = queries.size();
  std::vector<uint64_t> kw_ptr(Q + 1, 0), ph_ptr(Q + 1, 0);
  std::vector<uint32_t> kw, ph;
  for (size_t i = 0; i < Q; ++i) {
    ids(queries[i].queryTokenised, kw);
    ids(queries[i].phraseTokenised, ph);
    kw_ptr[i + 1] = kw.size();
    ph_ptr[i + 1] = ph.size();
  }
  std::vector<uint32_t> docs(Q * k), count(Q);
  std::vector<double> fin(Q * k), pr(Q * k);
  // topicProbs is a nil map in the shipped code (main_retrieve.go:87-88) => NULL => sqd = 0
  if (Q)
    must(ss_score_batch(e_, Q, kw_ptr.data(), ptr_or_null(kw), ph_ptr.data(), ptr_or_null(ph), nullptr, 0, k,
                        docs.data(), fin.data(), pr.data(), count.data()),
         "ss_score_batch");
  std::vector<std::vector<Rank_combined>> out(Q);
  for (size_t i = 0; i < Q; ++i) {
    out[i].resize(count[i]);
    for (uint32_t j = 0; j < count[i]; ++j) {
      out[i][j].DocHash = doc_keys_[docs[i * k + j]];
      out[i][j].PageRank = pr[i * k + j];
      out[i][j].FinalRank = fin[i * k + j];
    }
  }
  return out;
}

std::vector<Rank_combined> Index::Retrieve(const std::vector<std::string>& queryTokenised,
                                           const std::vector<std::string>& phraseTokenised, uint32_t k) const {
  return RetrieveBatch({Query{queryTokenised, phraseTokenised}}, k)[0];
}

struct BatchingRetriever::Impl {
  struct Pending {
    Index::Query q;
    std::promise<std::vector<Rank_combined>> done;
  };
  const Index& index;
  uint32_t k, max_batch, window_us;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Pending> queue;
  bool stop = false;
  uint64_t n_batches = 0, n_queries = 0;
  std::thread worker;
  Impl(const Index& ix, uint32_t k_, uint32_t mb, uint32_t win) : index(ix), k(k_), max_batch(mb), window_us(win) {
    worker = std::thread([this] { run(); });
  }
  void run() {
    for (;;) {
      std::vector<Pending> batch;
      {
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&] { return stop || !queue.empty(); });
        if (stop && queue.empty()) return;
        // let the window fill: concurrent callers that arrive within it share one kernel launch
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us);
        cv.wait_until(lock, deadline, [&] { return stop || queue.size() >= max_batch; });
        while (!queue.empty() && batch.size() < max_batch) {
          batch.push_back(std::move(queue.front()));
          queue.pop_front();
        }
      }
      std::vector<Index::Query> qs;
      qs.reserve(batch.size());
      for (auto& b : batch) qs.push_back(b.q);
      try {
        auto res = index.RetrieveBatch(qs, k);
        for (size_t i = 0; i < batch.size(); ++i) batch[i].done.set_value(std::move(res[i]));
      } catch (...) {
        for (auto& b : batch) b.done.set_exception(std::current_exception());
      }
      std::lock_guard<std::mutex> lock(mu);
      ++n_batches;
      n_queries += batch.size();
    }
  }
};

BatchingRetriever::BatchingRetriever(const Index& index, uint32_t k, uint32_t max_batch, uint32_t window_us)
    : impl_(new Impl(index, k, std::max(1u, max_batch), window_us)) {}
BatchingRetriever::~BatchingRetriever() {
  {
    std::lock_guard<std::mutex> lock(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv.notify_all();
  impl_->worker.join();
  delete impl_;
}
std::vector<Rank_combined> BatchingRetriever::Submit(const std::vector<std::string>& queryTokenised,
                                                     const std::vector<std::string>& phraseTokenised) {
  // The engine serves up to 256 keyword tokens per query (a longer phrase is simply never a hit, as in the
  // reference); a request beyond that fails HERE, for its own caller only, instead of failing the whole
  // coalesced batch for everybody who happened to share the window.
  if (queryTokenised.size() > 256)
    throw std::runtime_error("Retrieve: " + std::to_string(queryTokenised.size()) +
                             " keyword tokens in one query (the engine serves up to 256)");
  std::future<std::vector<Rank_combined>> fut;
  {
    std::lock_guard<std::mutex> lock(impl_->mu);
    impl_->queue.push_back({Index::Query{queryTokenised, phraseTokenised}, {}});
    fut = impl_->queue.back().done.get_future();
  }
  impl_->cv.notify_all();
  return fut.get();  // rethrows the worker's exception: the reference panics
}
uint64_t BatchingRetriever::batches() const {
  std::lock_guard<std::mutex> lock(impl_->mu);
  return impl_->n_batches;
}
uint64_t BatchingRetriever::queries() const {
  std::lock_guard<std::mutex> lock(impl_->mu);
  return impl_->n_queries;
}

std::vector<Rank_combined> Retrieve(ss_engine* e, const std::vector<std::string>& queryTokenised,
                                    const std::vector<std::string>& phraseTokenised, db::Table forw[6],
                                    db::Table inv[3]) {
  return Index(e, forw, inv).Retrieve(queryTokenised, phraseTokenised, 50);  // main_retrieve.go:99
}

}  // namespace retrieval

// ---- C entry points for tests and tools (ctypes) -------------------------------------------------
namespace {
thread_local std::string g_err;
db::Table* table_of(db::DB* d, const std::string& name) {
  if (name.size() == 4 && name.compare(0, 3, "inv") == 0 && name[3] >= '0' && name[3] <= '2') return &d->inv[name[3] - '0'];
  if (name.size() == 5 && name.compare(0, 4, "forw") == 0 && name[4] >= '0' && name[4] <= '5') return &d->forw[name[4] - '0'];
  fail("unknown table " + name);
}
std::vector<std::string> split(const char* s) {
  std::vector<std::string> out;
  std::istringstream in(s ? s : "");
  std::string tok;
  while (in >> tok) out.push_back(tok);
  return out;
}
template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
}  // namespace

extern "C" {
#define SSH_API __attribute__((visibility("default")))
SSH_API const char* ssh_last_error() { return g_err.c_str(); }
SSH_API void* ssh_db_new() { return new db::DB(); }
SSH_API void ssh_db_free(void* d) { delete (db::DB*)d; }
SSH_API int ssh_db_load_jsonl(void* d, const char* table, const char* path) {
  return guarded([&] { table_of((db::DB*)d, table)->LoadJsonl(path); });
}
SSH_API int ssh_db_save_jsonl(void* d, const char* table, const char* path) {
  return guarded([&] { table_of((db::DB*)d, table)->SaveJsonl(path); });
}
SSH_API long long ssh_db_rows(void* d, const char* table) {
  long long n = -1;
  guarded([&] { n = (long long)table_of((db::DB*)d, table)->size(); });
  return n;
}
// CSR export of forw[2] as a binary snapshot: u64 N, u64 E, row_ptr[N+1] u64, col_idx[E] u32, then N keys
// (32 bytes each when all keys are 32-char hashes, else newline separated)
SSH_API int ssh_export_graph(void* d, const char* path) {
  return guarded([&] {
    auto g = ranking::ExportGraph(((db::DB*)d)->forw[2]);
    std::ofstream out(path, std::ios::binary);
    if (!out) fail(std::string("cannot write ") + path);
    uint64_t n = g.keys.size(), e = g.col_idx.size();
    out.write((const char*)&n, 8);
    out.write((const char*)&e, 8);
    out.write((const char*)g.row_ptr.data(), (n + 1) * 8);
    out.write((const char*)g.col_idx.data(), e * 4);
    for (auto& k : g.keys) out << k << "\n";
  });
}
SSH_API int ssh_update_pagerank(void* d, void* engine, double damping, double eps) {
  return guarded([&] { ranking::UpdateTopicSensitivePagerank((ss_engine*)engine, damping, eps, ((db::DB*)d)->forw); });
}
SSH_API int ssh_update_term_weights(void* d, void* engine, const char* info) {
  return guarded([&] {
    db::DB* x = (db::DB*)d;
    ranking::UpdateTermWeights((ss_engine*)engine, std::string(info) == "title" ? &x->inv[0] : &x->inv[1], x->forw, info);
  });
}
// Serve `n` queries (newline separated; keyword hashes, then '|' and phrase hashes) from `threads`
// concurrent callers through a BatchingRetriever; out = one JSON array per line in query order.
// stats[0] = ss_score_batch calls made, stats[1] = queries served.
SSH_API int ssh_retrieve_concurrent(void* d, void* engine, const char* queries, int threads, int window_us,
                                    char* out, size_t cap, unsigned long long* stats) {
  return guarded([&] {
    db::DB* x = (db::DB*)d;
    std::vector<std::pair<std::vector<std::string>, std::vector<std::string>>> qs;
    std::istringstream in(queries ? queries : "");
    std::string line;
    while (std::getline(in, line)) {
      const size_t bar = line.find('|');
      qs.emplace_back(split(line.substr(0, bar).c_str()),
                      bar == std::string::npos ? std::vector<std::string>() : split(line.substr(bar + 1).c_str()));
    }
    retrieval::Index index((ss_engine*)engine, x->forw, x->inv);
    std::vector<std::vector<retrieval::Rank_combined>> res(qs.size());
    {
      retrieval::BatchingRetriever batcher(index, 50, 4096, (uint32_t)window_us);
      std::vector<std::thread> pool;
      std::atomic<size_t> next{0};
      std::string err;
      std::mutex err_mu;
      for (int t = 0; t < std::max(1, threads); ++t)
        pool.emplace_back([&] {
          for (size_t i = next++; i < qs.size(); i = next++) {
            try {
              res[i] = batcher.Submit(qs[i].first, qs[i].second);
            } catch (const std::exception& ex) {
              std::lock_guard<std::mutex> lock(err_mu);
              err = ex.what();
            }
          }
        });
      for (auto& th : pool) th.join();
      if (stats) {
        stats[0] = batcher.batches();
        stats[1] = batcher.queries();
      }
      if (!err.empty()) fail(err);
    }
    std::string js;
    for (auto& r : res) {
      js += "[";
      for (size_t i = 0; i < r.size(); ++i) {
        char buf[160];
        snprintf(buf, sizeof(buf), "%s{\"DocHash\": \"%s\", \"PageRank\": %.17g, \"FinalRank\": %.17g}", i ? ", " : "",
                 r[i].DocHash.c_str(), r[i].PageRank, r[i].FinalRank);
        js += buf;
      }
      js += "]\n";
    }
    if (js.size() + 1 > cap) fail("result buffer too small");
    memcpy(out, js.c_str(), js.size() + 1);
  });
}
// results as JSON: [{"DocHash": "...", "PageRank": x, "FinalRank": y}, ...] into out (NUL terminated)
SSH_API int ssh_retrieve(void* d, void* engine, const char* kw_hashes, const char* ph_hashes, char* out, size_t cap) {
  return guarded([&] {
    db::DB* x = (db::DB*)d;
    auto res = retrieval::Retrieve((ss_engine*)engine, split(kw_hashes), split(ph_hashes), x->forw, x->inv);
    std::string js = "[";
    for (size_t i = 0; i < res.size(); ++i) {
      char buf[160];
      snprintf(buf, sizeof(buf), "%s{\"DocHash\": \"%s\", \"PageRank\": %.17g, \"FinalRank\": %.17g}", i ? ", " : "",
               res[i].DocHash.c_str(), res[i].PageRank, res[i].FinalRank);
      js += buf;
    }
    js += "]";
    if (js.size() + 1 > cap) fail("result buffer too small");
    memcpy(out, js.c_str(), js.size() + 1);
  });
}
}
