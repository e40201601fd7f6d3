"""Synthetic CSR generation on the GPU with torch (same recipe as _synth.synthetic_csr, different RNG);
used by bench.py so that the 500k-cell configuration is ready in seconds."""
from __future__ import annotations

import torch


def synthetic_csr_torch(n_cells, n_genes, n_clusters, density=0.08, seed_data=1337, seed_labels=1338,
                        cluster_effect=True, device="cuda", chunk=16384):
    """Returns (indptr int32 [N+1], indices int32 [nnz], data float32 [nnz], labels int32 [N]) on `device`."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed_data)
    gl = torch.Generator(device=device)
    gl.manual_seed(seed_labels)
    labels = torch.randint(0, n_clusters, (n_cells,), generator=gl, device=device, dtype=torch.int32)
    conc1 = torch.full((n_genes,), 0.6, device=device, dtype=torch.float64)
    conc0 = torch.full((n_genes,), 0.6 * (1.0 - density) / density, device=device, dtype=torch.float64)
    # Beta via two Gammas (torch._standard_gamma has no generator arg on all builds -> seed globally)
    torch.manual_seed(seed_data)
    ga = torch._standard_gamma(conc1)
    gb = torch._standard_gamma(conc0)
    d_g = (ga / (ga + gb)).float()
    if cluster_effect:
        eff = torch.exp(0.5 * torch.randn((n_clusters, n_genes), generator=gen, device=device))
        d_cg = torch.clamp(d_g[None, :] * eff, 0.0, 1.0)
    else:
        d_cg = d_g[None, :].expand(n_clusters, n_genes).contiguous()
    counts, idx_parts, val_parts = [], [], []
    for s in range(0, n_cells, chunk):
        e = min(n_cells, s + chunk)
        u = torch.rand((e - s, n_genes), generator=gen, device=device)
        mask = u < d_cg[labels[s:e].long()]
        counts.append(mask.sum(dim=1).to(torch.int64))
        nz = mask.nonzero(as_tuple=False)
        idx_parts.append(nz[:, 1].to(torch.int32))
        k = nz.shape[0]
        gam = torch._standard_gamma(torch.full((k,), 1.5, device=device)) * 2.0
        val_parts.append(torch.log1p(gam).float())
    cnt = torch.cat(counts)
    indptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(cnt, 0)
    assert int(indptr[-1]) < 2 ** 31
    return indptr.to(torch.int32), torch.cat(idx_parts), torch.cat(val_parts), labels
