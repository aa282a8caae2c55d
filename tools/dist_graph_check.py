"""N-rank check of the sharded graph path (run on a GPU box):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_graph_check.py
Every rank builds the same seeded matrix, keeps its row shard, and the sharded results (PCA all-reduce, coordinate
all-gather, kNN of the local queries, replicated Leiden, cnv_score all-reduce) are compared with a purely local run."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, pandas as pd, scipy.sparse as sp, torch, torch.distributed as dist
import infercnvpy_b200 as cnv
from infercnvpy_b200.pp._neighbors import knn_device, neighbors_device, allgather_rows, fuzzy_graph_device, symmetrize
from infercnvpy_b200.tl._pca import pca_device

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(0)
n, K, n_clu = 7013, 640, 5
lab = rng.integers(0, n_clu, n)
centers = (rng.random((n_clu, K)) < 0.1) * rng.normal(size=(n_clu, K)) * 0.2
X = (centers[lab] + rng.normal(size=(n, K)) * 0.05 * (rng.random((n, K)) < 0.2)).astype(np.float32)
r0, r1 = cnv.shard_rows(n, 500, rank, world)
Xd_full, Xd = torch.from_numpy(X).to(dev), torch.from_numpy(X[r0:r1]).to(dev)
# PCA: K x K Gram all-reduce
Y, V, sv = pca_device(Xd, 20)
Yref, Vref, svref = pca_device(Xd_full, 20, reduce=False)
np.testing.assert_allclose(sv.cpu().numpy(), svref.cpu().numpy(), rtol=1e-9)
np.testing.assert_allclose(Y.cpu().numpy(), Yref[r0:r1].cpu().numpy(), rtol=1e-4, atol=1e-5)
# neighbours: all-gather of the coordinates, local queries; identical to the single-process lists
g = neighbors_device(Yref[r0:r1].contiguous(), 15)
assert g["row0"] == r0 and g["n_total"] == n
idx_ref, dist_ref = knn_device(Yref, 15)
assert torch.equal(g["idx"], idx_ref[r0:r1]) and torch.equal(g["dist"], dist_ref[r0:r1])
r_, c_, w_ = g["coo"]
A_sh = sp.csr_matrix((w_.cpu().numpy(), (r_.cpu().numpy(), c_.cpu().numpy())), shape=(n, n))
assert abs(A_sh - A_sh.T).max() < 1e-6 and A_sh.nnz > 14 * n
# public API on the shard: pca -> neighbors -> leiden -> cnv_score
obs = pd.DataFrame(index=[f"c{i}" for i in range(r0, r1)])
ad = cnv.AnnData(X[r0:r1], obs=obs)
ad.obsm["X_cnv"] = sp.csr_matrix(X[r0:r1])
cnv.tl.pca(ad, n_comps=20)
cnv.pp.neighbors(ad)
assert ad.obsp["cnv_neighbors_connectivities"].shape == (r1 - r0, n) and ad.uns["cnv_neighbors"]["shard"]["row0"] == r0
cnv.tl.leiden(ad)
cnv.tl.cnv_score(ad)
codes = torch.from_numpy(ad.obs["cnv_leiden"].cat.codes.values.astype(np.int64)).to(dev)
all_codes, _ = allgather_rows(codes)
all_codes = all_codes.cpu().numpy()
from sklearn.metrics import adjusted_rand_score
ari = adjusted_rand_score(lab, all_codes)
assert ari > 0.95, ari
# cnv_score of a cluster = mean |x| over ALL its cells (all ranks): check against numpy on the full matrix
sc = ad.obs["cnv_score"].values
for c in np.unique(all_codes[r0:r1]):
    want = np.abs(X[all_codes == c]).mean()
    got = sc[all_codes[r0:r1] == c][0]
    assert abs(got - want) <= 1e-6 * want, (c, got, want)
dist.barrier()
if rank == 0:
    print(f"DIST_OK world={world} n={n} clusters={len(np.unique(all_codes))} ARI_vs_planted={ari:.4f}")
dist.destroy_process_group()
