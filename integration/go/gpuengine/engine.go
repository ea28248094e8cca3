// Package gpuengine is the cgo binding of libspaghetti_gpu.so (include/spaghetti.h) plus the pieces the
// three drop-in files share: the process-wide engine, the hash -> dense-id dictionaries and their
// persistence between the offline pass (cmd/crawl) and the server (cmd/server).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain there).  Every C entry point used below is
// exercised through the identical ctypes binding (spaghettisearch_b200/capi.py) by the -m gpu tests, and
// the same export -> C ABI -> write-back flow runs, compiled, in csrc/host/host_mirror.cpp.
//
// Drop-in: copy integration/go/{gpuengine,ranking,retrieval} over the reference tree; the exported
// signatures of ranking.UpdateTopicSensitivePagerank, ranking.UpdateTermWeights and retrieval.Retrieve
// are unchanged, so cmd/crawl/start_crawl.go:175-177, cmd/server/server.go:47 and
// cmd/debug_retrieval.go:43 compile as they are.
package gpuengine

/*
#cgo CFLAGS: -I${SRCDIR}/../third_party/spaghetti/include
#cgo LDFLAGS: -L${SRCDIR}/../third_party/spaghetti -lspaghetti_gpu -Wl,-rpath,${SRCDIR}/../third_party/spaghetti
#include <stdlib.h>
#include "spaghetti.h"
*/
import "C"

import (
	"bufio"
	"errors"
	"os"
	"path/filepath"
	"sort"
	"sync"
	"unsafe"
)

// Table ids of ss_index_load / ss_term_weights (inv[0], inv[1]).
const (
	Title = int(C.SS_TITLE)
	Body  = int(C.SS_BODY)
)

// DictDir is where the sorted key files live (next to ./db_data/, database/database.go:99).
var DictDir = "./db_data/gpu_dict/"

var (
	once   sync.Once
	engine *C.ss_engine
)

// Engine returns the process-wide engine, creating it on first use.  There is no CPU fallback: without an
// sm_100 device ss_create fails and this panics, which is the reference's error convention
// (ranking/pagerank.go:20,29; retrieval/get_metadata.go:33,47).
func Engine() *C.ss_engine {
	once.Do(func() {
		cfg := C.ss_config{device: 0, flags: 0}
		must(C.ss_create(&cfg, &engine))
	})
	return engine
}

// must panics with the engine's message when an entry point failed (negative status).
func must(rc C.int) {
	if rc < 0 {
		panic(errors.New(C.GoString(C.ss_last_error())))
	}
}

// ---- pointer helpers (cgo: C must not retain Go pointers; every entry point copies before returning) ----

func u64(s []uint64) *C.uint64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint64_t)(unsafe.Pointer(&s[0]))
}
func u32(s []uint32) *C.uint32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&s[0]))
}
func f32(s []float32) *C.float {
	if len(s) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&s[0]))
}
func f64(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}
func i64(s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}

// ---- Go-typed wrappers of the C ABI (cgo types are package local, so ranking/ and retrieval/ call these) ----

// LoadGraph = ss_graph_load_csr: out-edge CSR of forw[2] on dense ids.
func LoadGraph(rowPtr []uint64, colIdx []uint32) {
	must(C.ss_graph_load_csr(Engine(), C.uint64_t(len(rowPtr)-1), C.uint64_t(len(colIdx)), u64(rowPtr), u32(colIdx)))
}

// Pagerank = ss_pagerank for up to 16 topics; returns rank [N][len(numPages)] row major.
func Pagerank(n int, damping, eps float64, numPages []int64) []float64 {
	out := make([]float64, n*len(numPages))
	must(C.ss_pagerank(Engine(), C.double(damping), C.double(eps), C.uint32_t(len(numPages)), i64(numPages),
		0 /* unbounded like the reference */, f64(out), nil))
	return out
}

// Postings is one inverted table in the engine's layout (spaghetti.h, ss_index_load).
type Postings struct {
	TermPtr []uint64  // [V+1]
	DocIDs  []uint32  // ascending within a term
	W       []float32 // listPos[0]
	PosPtr  []uint64  // [P+1]
	Pos     []float32 // listPos[1:]
}

// IndexLoad = ss_index_load.
func IndexLoad(table int, nDocs int, p *Postings) {
	must(C.ss_index_load(Engine(), C.int(table), C.uint64_t(len(p.TermPtr)-1), C.uint64_t(nDocs), u64(p.TermPtr),
		u32(p.DocIDs), f32(p.W), u64(p.PosPtr), f32(p.Pos)))
}

// TermWeights = ss_term_weights: returns the weighted listPos[0] of every posting and sqrt(sum w^2) per doc.
func TermWeights(table int, totalDocs float64, nPostings, nDocs int) (w []float32, mag []float64) {
	w, mag = make([]float32, nPostings), make([]float64, nDocs)
	must(C.ss_term_weights(Engine(), C.int(table), C.double(totalDocs), nil, f32(w), f64(mag)))
	return
}

// SetDocNorms = ss_set_doc_norms (forw[4] column `info`), SetPagerank = ss_set_pagerank (forw[3] rows).
func SetDocNorms(table int, mag []float64) {
	must(C.ss_set_doc_norms(Engine(), C.int(table), C.uint64_t(len(mag)), f64(mag)))
}
func SetPagerank(nDocs, nTopics int, rank []float64) {
	must(C.ss_set_pagerank(Engine(), C.uint64_t(nDocs), C.uint32_t(nTopics), f64(rank)))
}

// Result is one row of ss_score_batch's output.
type Result struct {
	Doc       uint32
	FinalRank float64
	PageRank  float64
}

// ScoreBatch = ss_score_batch for a batch of queries given as token-id lists; topicProbs nil = the shipped
// behaviour (retrieval/main_retrieve.go:87-88: nil map => sqd = 0).
func ScoreBatch(kw [][]uint32, ph [][]uint32, topicProbs []float64, k int) [][]Result {
	nq := len(kw)
	kwPtr, phPtr := make([]uint64, nq+1), make([]uint64, nq+1)
	var kwT, phT []uint32
	for q := 0; q < nq; q++ {
		kwT = append(kwT, kw[q]...)
		phT = append(phT, ph[q]...)
		kwPtr[q+1], phPtr[q+1] = uint64(len(kwT)), uint64(len(phT))
	}
	docs, cnt := make([]uint32, nq*k), make([]uint32, nq)
	fin, pr := make([]float64, nq*k), make([]float64, nq*k)
	must(C.ss_score_batch(Engine(), C.uint64_t(nq), u64(kwPtr), u32(kwT), u64(phPtr), u32(phT), f64(topicProbs), 0,
		C.uint32_t(k), u32(docs), f64(fin), f64(pr), u32(cnt)))
	out := make([][]Result, nq)
	for q := 0; q < nq; q++ {
		for j := 0; j < int(cnt[q]); j++ {
			out[q] = append(out[q], Result{docs[q*k+j], fin[q*k+j], pr[q*k+j]})
		}
	}
	return out
}

// ---- dictionaries -----------------------------------------------------------------------------------
//
// Dense id = rank of the 32-hex md5 key in ascending order (include/spaghetti.h "Dense ids"), so the engine's
// tie-break "ascending doc id" is "ascending doc hash" and is the same in every process that builds the
// dictionary from the same key set.

// Dict maps 32-hex md5 keys to dense ids and back.
type Dict struct {
	Keys []string          // id -> key, ascending
	ID   map[string]uint32 // key -> id
}

// NewDict sorts the key set and numbers it.
func NewDict(keys []string) *Dict {
	sort.Strings(keys)
	d := &Dict{Keys: keys, ID: make(map[string]uint32, len(keys))}
	for i, k := range keys {
		d.ID[k] = uint32(i)
	}
	return d
}

// Lookup returns the dense id, or 0xFFFFFFFF for a key that is not in the dictionary: the engine treats a
// term id >= n_terms as an unknown term with an empty posting list, which mirrors the ErrKeyNotFound
// tolerance of retrieval/main_retrieve.go:193,218.
func (d *Dict) Lookup(key string) uint32 {
	if id, ok := d.ID[key]; ok {
		return id
	}
	return 0xFFFFFFFF
}

// Save writes one key per line; the line number is the id.
func (d *Dict) Save(name string) {
	if err := os.MkdirAll(DictDir, 0755); err != nil {
		panic(err)
	}
	f, err := os.Create(filepath.Join(DictDir, name))
	if err != nil {
		panic(err)
	}
	defer f.Close()
	w := bufio.NewWriter(f)
	for _, k := range d.Keys {
		if _, err = w.WriteString(k + "\n"); err != nil {
			panic(err)
		}
	}
	if err = w.Flush(); err != nil {
		panic(err)
	}
}

// LoadDict reads a file written by Save; nil if it does not exist (offline pass not run yet).
func LoadDict(name string) *Dict {
	f, err := os.Open(filepath.Join(DictDir, name))
	if err != nil {
		return nil
	}
	defer f.Close()
	var keys []string
	sc := bufio.NewScanner(f)
	for sc.Scan() {
		keys = append(keys, sc.Text())
	}
	if err = sc.Err(); err != nil {
		panic(err)
	}
	d := &Dict{Keys: keys, ID: make(map[string]uint32, len(keys))}
	for i, k := range keys {
		d.ID[k] = uint32(i)
	}
	return d
}

// Names of the persisted dictionaries.
const (
	DocDictFile  = "docs.keys"  // node set of forw[2] = doc id space (ranking/pagerank.go:24-44)
	TermDictFile = "terms.keys" // union of the term keys of inv[0] and inv[1]
	TopicFile    = "topics.keys"
)
