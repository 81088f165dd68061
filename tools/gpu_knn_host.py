import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, jsa_rag_b200 as eng
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
n = 4_125_000
passages = [{"id": str(i), "title": f"title {i}", "text": "lorem ipsum " * 20} for i in range(n)]
idx = eng.B200Index(); idx.init_embeddings(passages, dim=768)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    idx.embeddings[:, s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half().T
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g, device=dev), dim=1)
for _ in range(3): idx.search_knn(q, 100)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): docs, scores = idx.search_knn(q, 100)
t_knn = (time.perf_counter() - t0) / 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): s_, i_ = idx.search(q, 100); torch.cuda.synchronize()
t_s = (time.perf_counter() - t0) / 20
print(f"search_knn (docs+scores lists) {t_knn*1e3:.3f} ms ; tensor search {t_s*1e3:.3f} ms ; host part {1e3*(t_knn-t_s):.3f} ms")
