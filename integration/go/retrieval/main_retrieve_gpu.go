// Drop-in replacement of the score/blend/top-k core of retrieval/main_retrieve.go: same package, same
// exported signature (retrieval/main_retrieve.go:15).  Query parsing (getPhrase, parser.Laundry, md5) and the
// hydration of the returned documents (getDocInfo, getSummary, resultFormat) are the reference's own
// functions; everything between them -- posting retrieval, phrase matching, aggregation, cosine,
// PageRank blend, ordering, truncation to 50 -- is one ss_score_batch call.
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain); see integration/go/gpuengine/engine.go.
package retrieval

import (
	"context"
	"crypto/md5"
	"encoding/hex"
	"encoding/json"
	"sort"
	"strings"
	"sync"
	"time"

	db "github.com/nwihardjo/SpaghettiSearch/database"
	gpu "github.com/nwihardjo/SpaghettiSearch/gpuengine"
	"github.com/nwihardjo/SpaghettiSearch/parser"
)

var (
	loadOnce sync.Once
	docDict  *gpu.Dict // dense doc id <-> doc hash
	termDict *gpu.Dict // dense term id <-> term hash (union of inv[0] and inv[1])
)

// LoadIndex puts the weighted index on the GPU once per process (cmd/server/server.go calls it after
// DB_init, or the first Retrieve does): inv[0|1] rows -> ss_index_load (weights are already applied by the
// offline pass, so ss_term_weights is NOT called again), forw[4] -> ss_set_doc_norms, forw[3] ->
// ss_set_pagerank.  The dictionaries are the sorted key files written by the offline pass.
func LoadIndex(ctx context.Context, forw []db.DB, inv []db.DB) {
	loadOnce.Do(func() {
		docDict = gpu.LoadDict(gpu.DocDictFile)
		if docDict == nil {
			panic("gpu_dict/docs.keys is missing: run the offline pass (cmd/crawl) first")
		}
		// term dictionary = union of the two tables' keys
		comps := make([]*db.Collector, 2)
		seen := make(map[string]struct{})
		for tb := 0; tb < 2; tb++ {
			c, err := inv[tb].Iterate(ctx)
			if err != nil {
				panic(err)
			}
			comps[tb] = c
			for _, kv := range c.KV {
				seen[string(kv.Key)] = struct{}{}
			}
		}
		terms := make([]string, 0, len(seen))
		for t := range seen {
			terms = append(terms, t)
		}
		termDict = gpu.NewDict(terms)
		termDict.Save(gpu.TermDictFile)
		for tb := 0; tb < 2; tb++ {
			gpu.IndexLoad(tb, len(docDict.Keys), exportWeighted(comps[tb]))
		}
		// forw[4]: {"title": m, "body": m}; a missing key reads as 0 like the Go map (get_metadata.go:57-58)
		magT, magB := make([]float64, len(docDict.Keys)), make([]float64, len(docDict.Keys))
		c, err := forw[4].Iterate(ctx)
		if err != nil {
			panic(err)
		}
		for _, kv := range c.KV {
			var val map[string]float64
			if err = json.Unmarshal(kv.Value, &val); err != nil {
				panic(err)
			}
			if id, ok := docDict.ID[string(kv.Key)]; ok {
				magT[id], magB[id] = val["title"], val["body"]
			}
		}
		gpu.SetDocNorms(gpu.Title, magT)
		gpu.SetDocNorms(gpu.Body, magB)
		// forw[3]: {category: rank}; topic order = the sorted category keys written by the offline pass
		if topics := gpu.LoadDict(gpu.TopicFile); topics != nil && len(topics.Keys) > 0 {
			T := len(topics.Keys)
			rank := make([]float64, len(docDict.Keys)*T)
			c, err = forw[3].Iterate(ctx)
			if err != nil {
				panic(err)
			}
			for _, kv := range c.KV {
				var val map[string]float64
				if err = json.Unmarshal(kv.Value, &val); err != nil {
					panic(err)
				}
				if id, ok := docDict.ID[string(kv.Key)]; ok {
					for t, name := range topics.Keys {
						rank[int(id)*T+t] = val[name]
					}
				}
			}
			gpu.SetPagerank(len(docDict.Keys), T, rank)
		}
	})
}

// exportWeighted lays one inverted table out term-major on the SHARED term dictionary (a term missing from
// this table is an empty row), docs ascending by dense id.
func exportWeighted(comp *db.Collector) *gpu.Postings {
	rows := make(map[uint32]map[string][]float32, len(comp.KV))
	for _, kv := range comp.KV {
		var val map[string][]float32
		if err := json.Unmarshal(kv.Value, &val); err != nil {
			panic(err)
		}
		rows[termDict.ID[string(kv.Key)]] = val
	}
	p := &gpu.Postings{TermPtr: make([]uint64, len(termDict.Keys)+1), PosPtr: []uint64{0}}
	for t := range termDict.Keys {
		val := rows[uint32(t)]
		hashes := make([]string, 0, len(val))
		for h := range val {
			if _, ok := docDict.ID[h]; ok {
				hashes = append(hashes, h)
			}
		}
		sort.Strings(hashes)
		for _, h := range hashes {
			listPos := val[h]
			p.DocIDs = append(p.DocIDs, docDict.ID[h])
			p.W = append(p.W, listPos[0])
			p.Pos = append(p.Pos, listPos[1:]...)
			p.PosPtr = append(p.PosPtr, uint64(len(p.Pos)))
		}
		p.TermPtr[t+1] = uint64(len(p.DocIDs))
	}
	return p
}

// termIDs maps laundered tokens to dense term ids through their md5 hex (main_retrieve.go:29-36); a term
// that is in neither table becomes 0xFFFFFFFF = empty posting list (main_retrieve.go:193,218).
func termIDs(tokens []string) []uint32 {
	ids := make([]uint32, len(tokens))
	for i, tok := range tokens {
		sum := md5.Sum([]byte(tok))
		ids[i] = termDict.Lookup(hex.EncodeToString(sum[:]))
	}
	return ids
}

// ---- request coalescing (SURVEY.md 8(f)-2; compiled twin: retrieval::BatchingRetriever in host_mirror.cpp) ----
// net/http runs one goroutine per request (cmd/server/server.go:32-52); the engine is built for batches, so
// concurrent Retrieve calls that arrive within batchWindow share one ss_score_batch.

type pending struct {
	kw, ph []uint32
	done   chan []gpu.Result
}

var (
	queue       = make(chan *pending, 4096)
	batcherOnce sync.Once
	batchWindow = 200 * time.Microsecond
	maxBatch    = 1024
)

func batcher() {
	for first := range queue {
		batch := []*pending{first}
		timer := time.NewTimer(batchWindow)
	collect:
		for len(batch) < maxBatch {
			select {
			case p := <-queue:
				batch = append(batch, p)
			case <-timer.C:
				break collect
			}
		}
		timer.Stop()
		kw, ph := make([][]uint32, len(batch)), make([][]uint32, len(batch))
		for i, p := range batch {
			kw[i], ph[i] = p.kw, p.ph
		}
		// topicProbs is a nil map in the shipped code (main_retrieve.go:87-88) => NULL => sqd = 0
		res := gpu.ScoreBatch(kw, ph, nil, 50) // keep the first 50 (main_retrieve.go:99-103)
		for i, p := range batch {
			p.done <- res[i]
		}
	}
}

// Retrieve keeps the reference's signature and result type (retrieval/main_retrieve.go:15).
func Retrieve(query string, ctx context.Context, forw []db.DB, inv []db.DB) []Rank_combined {
	LoadIndex(ctx, forw, inv)
	batcherOnce.Do(func() { go batcher() })

	// ---- query parsing: unchanged (main_retrieve.go:19-36)
	phrases := getPhrase(query)
	for _, term := range phrases {
		query = strings.Replace(query, "\""+string(term)+"\"", "", 1)
	}
	queryTokenised := parser.Laundry(strings.Join(strings.Fields(query), " "))
	phraseTokenised := parser.Laundry(strings.Join(phrases, " ")) // all phrases form ONE phrase (:26)

	// ---- score / blend / top-50 on the GPU
	p := &pending{kw: termIDs(queryTokenised), ph: termIDs(phraseTokenised), done: make(chan []gpu.Result, 1)}
	queue <- p
	top := <-p.done

	// ---- hydrate ONLY the returned documents (the reference does it for every match, get_metadata.go:27-28)
	out := make([]Rank_combined, len(top))
	var wg sync.WaitGroup
	for i, r := range top {
		wg.Add(1)
		go func(i int, r gpu.Result) {
			defer wg.Done()
			hash := docDict.Keys[r.Doc]
			meta := getDocInfo(ctx, hash, forw)         // get_metadata.go:211, returns <-chan Rank_combined
			summary := getSummary(hash, query, phrases) // get_metadata.go:79
			m := <-meta
			m.PageRank, m.FinalRank, m.Summary = r.PageRank, r.FinalRank, <-summary
			out[i] = m
		}(i, r)
	}
	wg.Wait()
	return out
}
