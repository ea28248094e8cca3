// Drop-in replacement of ranking/term_weighting.go: same package, same exported signature
// (ranking/term_weighting.go:10).  Export of the inverted table, write-back of the weighted postings and
// the forw[4] merge (saveMagnitude's rules, term_weighting.go:59-123) stay in Go; idf, the fp32 weights and
// the doc norms are computed by libspaghetti_gpu.so (bit-identical to the reference's arithmetic, see
// tests/test_scoring_gpu.py::test_idf_matches_go_log2_bitwise).
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain); see integration/go/gpuengine/engine.go.
package ranking

import (
	"context"
	"encoding/json"
	"sort"

	db "github.com/nwihardjo/SpaghettiSearch/database"
	gpu "github.com/nwihardjo/SpaghettiSearch/gpuengine"
)

// exported is one inverted table decoded into the engine's layout, remembering the Badger keys.
type exported struct {
	terms []string   // term hash of row t (ascending = dense term id inside this table)
	docs  [][]string // doc hashes of row t, ascending doc id
	post  gpu.Postings
}

// exportPostings decodes every row of inv (term_weighting.go:29-35) into term-major arrays:
// listPos[0] -> W (normTF), listPos[1:] -> Pos; docs of a row sorted by dense doc id.
func exportPostings(comp *db.Collector, docDict *gpu.Dict) *exported {
	type row struct {
		term string
		val  map[string][]float32
	}
	rows := make([]row, len(comp.KV))
	for i, kv := range comp.KV {
		var val map[string][]float32
		if err := json.Unmarshal(kv.Value, &val); err != nil {
			panic(err)
		}
		rows[i] = row{string(kv.Key), val}
	}
	sort.Slice(rows, func(a, b int) bool { return rows[a].term < rows[b].term })
	ex := &exported{terms: make([]string, len(rows)), docs: make([][]string, len(rows))}
	ex.post.TermPtr = make([]uint64, len(rows)+1)
	ex.post.PosPtr = []uint64{0}
	for t, r := range rows {
		ex.terms[t] = r.term
		hashes := make([]string, 0, len(r.val))
		for h := range r.val {
			hashes = append(hashes, h)
		}
		sort.Strings(hashes) // ascending hash == ascending dense id
		ex.docs[t] = hashes
		for _, h := range hashes {
			listPos := r.val[h]
			id, ok := docDict.ID[h]
			if !ok { // a doc that is not a node of forw[2]: the reference would panic later in computeFinalRank
				panic("document " + h + " of the inverted table is missing from forw[2]/forw[3]")
			}
			ex.post.DocIDs = append(ex.post.DocIDs, id)
			ex.post.W = append(ex.post.W, listPos[0])
			ex.post.Pos = append(ex.post.Pos, listPos[1:]...)
			ex.post.PosPtr = append(ex.post.PosPtr, uint64(len(ex.post.Pos)))
		}
		ex.post.TermPtr[t+1] = uint64(len(ex.post.DocIDs))
	}
	return ex
}

// writeBackWeights stores listPos[0] = w for every posting (term_weighting.go:42-49): one BatchSet per row.
func writeBackWeights(ctx context.Context, inv *db.DB, ex *exported, w []float32) {
	bw := (*inv).BatchWrite_init(ctx)
	defer bw.Cancel(ctx)
	for t, term := range ex.terms {
		val := make(map[string][]float32, len(ex.docs[t]))
		for j, h := range ex.docs[t] {
			x := ex.post.TermPtr[t] + uint64(j)
			listPos := make([]float32, 0, 1+ex.post.PosPtr[x+1]-ex.post.PosPtr[x])
			listPos = append(listPos, w[x])
			listPos = append(listPos, ex.post.Pos[ex.post.PosPtr[x]:ex.post.PosPtr[x+1]]...)
			val[h] = listPos
		}
		if err := bw.BatchSet(ctx, term, val); err != nil {
			panic(err)
		}
	}
	if err := bw.Flush(ctx); err != nil {
		panic(err)
	}
}

// saveMagnitudeDense applies saveMagnitude's rules (term_weighting.go:59-123) to the dense norm vector:
// only docs with a posting in THIS table are candidates (the reference's pageMagnitude map holds exactly
// those); existing forw[4] rows get info merged in -- with sqrt(0) = 0 when the doc has no posting here
// (term_weighting.go:97) --, the remaining candidates get new rows {info: norm}.
func saveMagnitudeDense(ctx context.Context, forw *db.DB, info string, docDict *gpu.Dict, ex *exported, mag []float64) {
	has := make([]bool, len(docDict.Keys))
	for _, id := range ex.post.DocIDs {
		has[id] = true
	}
	comp, err := (*forw).Iterate(ctx)
	if err != nil {
		panic(err)
	}
	bw := (*forw).BatchWrite_init(ctx)
	defer bw.Cancel(ctx)
	done := make([]bool, len(docDict.Keys))
	for _, kv := range comp.KV {
		key := string(kv.Key)
		var val map[string]float64
		if err = json.Unmarshal(kv.Value, &val); err != nil {
			panic(err)
		}
		m := 0.0
		if id, ok := docDict.ID[key]; ok {
			done[id] = true
			if has[id] {
				m = mag[id]
			}
		}
		val[info] = m
		if err = bw.BatchSet(ctx, key, val); err != nil {
			panic(err)
		}
	}
	for id, key := range docDict.Keys {
		if has[id] && !done[id] {
			if err = bw.BatchSet(ctx, key, map[string]float64{info: mag[id]}); err != nil {
				panic(err)
			}
		}
	}
	if err = bw.Flush(ctx); err != nil {
		panic(err)
	}
}

// UpdateTermWeights keeps the reference's signature and side effects (term_weighting.go:10-57).
func UpdateTermWeights(ctx context.Context, inv *db.DB, forw []db.DB, info string) {
	nodesCompressed, err := forw[3].Iterate(ctx) // totalDocs = rows of forw[3] (term_weighting.go:12-17)
	if err != nil {
		panic(err)
	}
	totalDocs := float64(len(nodesCompressed.KV))
	docDict := gpu.LoadDict(gpu.DocDictFile) // written by UpdateTopicSensitivePagerank, which runs first
	if docDict == nil {                     // (cmd/crawl/start_crawl.go:175-177); rebuild it from forw[3] otherwise
		keys := make([]string, len(nodesCompressed.KV))
		for i, kv := range nodesCompressed.KV {
			keys[i] = string(kv.Key)
		}
		docDict = gpu.NewDict(keys)
		docDict.Save(gpu.DocDictFile)
	}
	comp, err := (*inv).Iterate(ctx)
	if err != nil {
		panic(err)
	}
	ex := exportPostings(comp, docDict)
	table := gpu.Body
	if info == "title" {
		table = gpu.Title
	}
	gpu.IndexLoad(table, len(docDict.Keys), &ex.post)
	w, mag := gpu.TermWeights(table, totalDocs, len(ex.post.DocIDs), len(docDict.Keys))
	writeBackWeights(ctx, inv, ex, w)
	saveMagnitudeDense(ctx, &forw[4], info, docDict, ex, mag)
}
