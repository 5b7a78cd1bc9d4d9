"""Binary ingest format for read counts.

The reference ingests `read_counts.tsv` into BEAST XML text (app/datacollector/DataCollector.java:213-364), which is
infeasible at 10^4 cells x 10^6 sites (10^10 tuples).  A shard file holds one contiguous site range of the data set in the
device's own order, ready to be memory-mapped and streamed to the GPU:

    bytes 0..4095     header: b"SIEVEB2\\n" + JSON (utf-8, zero padded) with n_cells, n_sites, site_begin, n_sites_total,
                      dtype ("uint16" | "int32"), n_background_sites, cell_names, size_factors (global, may be null)
    bytes 4096..      tensor [cell][site][4] = (alt1, alt2, alt3, coverage), little endian

Cells are sorted by name and sites by locus exactly as DataCollector writes them (:477-502, :541-583); the site split over
shards is calcPatternPoints (likelihood/ThreadedScsTreeLikelihood.java:227-238).  Size factors are global
(computed over all sites before the split, :119-131) and stored in every shard.
"""
import json

import numpy as np

MAGIC = b"SIEVEB2\n"
HEADER_BYTES = 4096


def _points(n_sites, n_shards):
    pts = [0]
    for i in range(n_shards):
        pts.append(pts[-1] + n_sites // n_shards + (1 if i < n_sites % n_shards else 0))
    return pts


def write_shard(path, counts_site_cell_4, site_begin, n_sites_total, cell_names=None, size_factors=None,
                n_background_sites=None, dtype=None):
    """counts_site_cell_4: int [sites of this shard][cell][4] (the loader's / alignment's order)."""
    c = np.asarray(counts_site_cell_4)
    if c.ndim != 3 or c.shape[2] != 4:
        raise ValueError("counts must be [site][cell][4]")
    if (c < 0).any():
        raise ValueError("Read counts should be non-negative integers")
    if dtype is None:
        dtype = "uint16" if c.max(initial=0) < 65536 else "int32"
    if dtype == "uint16" and c.max(initial=0) >= 65536:
        raise ValueError("coverage does not fit uint16")
    S, n = c.shape[:2]
    meta = dict(n_cells=int(n), n_sites=int(S), site_begin=int(site_begin), n_sites_total=int(n_sites_total), dtype=dtype,
                n_background_sites=None if n_background_sites is None else int(n_background_sites),
                cell_names=None if cell_names is None else list(cell_names),
                size_factors=None if size_factors is None else [float(x) for x in size_factors])
    blob = MAGIC + json.dumps(meta).encode("utf-8")
    if len(blob) > HEADER_BYTES:
        # long cell-name lists go to a side file so that the tensor offset stays fixed
        with open(path + ".cells", "w") as f:
            f.write("\n".join(meta["cell_names"] or []))
        meta["cell_names"] = None
        blob = MAGIC + json.dumps(meta).encode("utf-8")
        if len(blob) > HEADER_BYTES:
            with open(path + ".sizefactors", "w") as f:
                f.write("\n".join(repr(x) for x in meta["size_factors"] or []))
            meta["size_factors"] = None
            blob = MAGIC + json.dumps(meta).encode("utf-8")
    with open(path, "wb") as f:
        f.write(blob.ljust(HEADER_BYTES, b"\0"))
        np.ascontiguousarray(c.transpose(1, 0, 2), dtype=np.dtype(dtype).newbyteorder("<")).tofile(f)
    return meta


def open_shard(path):
    """-> (metadata dict, np.memmap [cell][site][4])"""
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
    if not head.startswith(MAGIC):
        raise ValueError("not a sieve_b200 shard file: " + path)
    meta = json.loads(head[len(MAGIC):].rstrip(b"\0").decode("utf-8"))
    import os
    if meta.get("cell_names") is None and os.path.exists(path + ".cells"):
        meta["cell_names"] = open(path + ".cells").read().split("\n")
    if meta.get("size_factors") is None and os.path.exists(path + ".sizefactors"):
        meta["size_factors"] = [float(x) for x in open(path + ".sizefactors").read().split("\n") if x]
    arr = np.memmap(path, dtype=np.dtype(meta["dtype"]).newbyteorder("<"), mode="r", offset=HEADER_BYTES,
                    shape=(meta["n_cells"], meta["n_sites"], 4))
    return meta, arr


def write_shards(prefix, counts_site_cell_4, n_shards, cell_names=None, size_factors=None, n_background_sites=None):
    """Splits the sites like calcPatternPoints and writes prefix.<r>.sieve for r in 0..n_shards-1."""
    c = np.asarray(counts_site_cell_4)
    pts = _points(c.shape[0], n_shards)
    paths = []
    for r in range(n_shards):
        p = "%s.%d.sieve" % (prefix, r)
        write_shard(p, c[pts[r]:pts[r + 1]], pts[r], c.shape[0], cell_names, size_factors, n_background_sites)
        paths.append(p)
    return paths


def convert_tsv(tsv_path, cell_names_path, prefix, n_shards=1, size_factors=None):
    """read_counts.tsv (+ cell names) -> shard files; size factors may be supplied (e.g. from
    ScsCudaLikelihoodCore.compute_size_factors on the whole data set)."""
    from .data import load_read_counts_tsv
    d = load_read_counts_tsv(tsv_path, cell_names_path)
    return write_shards(prefix, d["counts"], n_shards, d["cell_names"], size_factors, d["n_background_sites"])
