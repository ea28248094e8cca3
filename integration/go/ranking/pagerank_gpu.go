// Drop-in replacement of ranking/pagerank.go: same package, same exported signature
// (ranking/pagerank.go:14), the power iteration runs in libspaghetti_gpu.so.  Badger iteration, JSON
// decode and the forw[3] write-back are the reference's own code paths.
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain); see integration/go/gpuengine/engine.go.
package ranking

import (
	"context"
	"encoding/json"

	db "github.com/nwihardjo/SpaghettiSearch/database"
	gpu "github.com/nwihardjo/SpaghettiSearch/gpuengine"
)

// UpdateTopicSensitivePagerank keeps the reference's signature and side effects:
// forw[3][node] = {category: rank} for every node of parents U children (ranking/pagerank.go:66-82).
func UpdateTopicSensitivePagerank(ctx context.Context, dampingFactor float64, convergenceCriterion float64, forward []db.DB) {
	// ---- export forw[2] (ranking/pagerank.go:18-44)
	nodesCompressed, err := forward[2].Iterate(ctx)
	if err != nil {
		panic(err)
	}
	children := make(map[string][]string, len(nodesCompressed.KV))
	set := make(map[string]struct{}, len(nodesCompressed.KV))
	for _, kv := range nodesCompressed.KV {
		var kids []string
		if err = json.Unmarshal(kv.Value, &kids); err != nil {
			panic(err)
		}
		key := string(kv.Key)
		children[key] = kids
		set[key] = struct{}{}
		for _, c := range kids {
			set[c] = struct{}{}
		}
	}
	keys := make([]string, 0, len(set))
	for k := range set {
		keys = append(keys, k)
	}
	docs := gpu.NewDict(keys) // dense id = rank of the hex hash
	docs.Save(gpu.DocDictFile)

	rowPtr := make([]uint64, len(docs.Keys)+1)
	colIdx := make([]uint32, 0, len(nodesCompressed.KV)*8)
	for i, k := range docs.Keys {
		for _, c := range children[k] { // each list entry counts (ranking/pagerank.go:136-142)
			colIdx = append(colIdx, docs.ID[c])
		}
		rowPtr[i+1] = uint64(len(colIdx))
	}

	// ---- forw[5] categories (ranking/pagerank.go:47-61); topic order = ascending category key
	cats, err := forward[5].Iterate(ctx)
	if err != nil {
		panic(err)
	}
	pages := make(map[string]int64, len(cats.KV))
	names := make([]string, 0, len(cats.KV))
	for _, kv := range cats.KV {
		val := make(map[string]float64, 2)
		if err = json.Unmarshal(kv.Value, &val); err != nil {
			panic(err)
		}
		names = append(names, string(kv.Key))
		pages[string(kv.Key)] = int64(int(val["numPages"])) // ranking/pagerank.go:61
	}
	topics := gpu.NewDict(names)
	topics.Save(gpu.TopicFile)
	numPages := make([]int64, len(topics.Keys))
	for t, name := range topics.Keys {
		numPages[t] = pages[name]
	}

	// ---- the hot path (ranking/pagerank.go:54-63,85-145)
	N, T := len(docs.Keys), len(topics.Keys)
	gpu.LoadGraph(rowPtr, colIdx)
	rank := make([]float64, N*T) // [N][T] row major
	for lo := 0; lo < T; lo += 16 { // topics are independent runs: slabs of 16 columns
		hi := lo + 16
		if hi > T {
			hi = T
		}
		w := hi - lo
		slab := gpu.Pagerank(N, dampingFactor, convergenceCriterion, numPages[lo:hi])
		for v := 0; v < N; v++ {
			copy(rank[v*T+lo:v*T+hi], slab[v*w:(v+1)*w])
		}
	}

	// ---- write forw[3] (ranking/pagerank.go:66-82); T == 0 writes {} for every node, as the reference does
	bw := forward[3].BatchWrite_init(ctx)
	defer bw.Cancel(ctx)
	for v, k := range docs.Keys {
		PR := make(map[string]float64, T)
		for t, name := range topics.Keys {
			PR[name] = rank[v*T+t]
		}
		if err = bw.BatchSet(ctx, k, PR); err != nil {
			panic(err)
		}
	}
	if err = bw.Flush(ctx); err != nil {
		panic(err)
	}
}
