// core_maxsum.cuh -- host side of the max-sum pass (kernels: maxsum.cuh): sieve_ml_trace and its getters.
// Included by sieve_core.cu only.
#pragma once
// ---- max-sum pass: maximum-likelihood genotypes for variant calling (maxsum.cuh) --------------------------------
// Post-order list of the whole current tree from the children the core has recorded (the planner's map).
static int sv_ml_build_ops(sieve_core* h, std::vector<SvMlOp>& ops) {
    const int n = h->n;
    if (h->root < 0) return sv_fail(SIEVE_ERR_STATE, "sieve_ml_trace: sieve_set_root has not been called");
    std::vector<std::pair<int, int>> st;
    st.push_back({h->root, 0});
    while (!st.empty()) {
        auto [v, state] = st.back();
        st.pop_back();
        if (v < n) continue;
        const int a = h->eph_c1[v], b = h->eph_c2[v];
        if (a == -2) return sv_fail(SIEVE_ERR_STATE, "sieve_ml_trace: the tree has a node that was never computed");
        if (state == 0) {
            st.push_back({v, 1});
            if (b >= 0) st.push_back({b, 0});
            st.push_back({a, 0});
        } else {
            SvMlOp op{};
            op.dst = v - n;
            op.n1 = a;
            op.c1 = a < n ? a : a - n;
            if (a < n) op.flags |= SV_OP_LEAF1;
            if (b < 0) {
                op.flags |= SV_OP_UNARY;
                op.c2 = -1;
                op.n2 = 0;
            } else {
                op.n2 = b;
                op.c2 = b < n ? b : b - n;
                if (b < n) op.flags |= SV_OP_LEAF2;
            }
            ops.push_back(op);
        }
    }
    return SIEVE_OK;
}

static int sv_ml_require(sieve_core* h, const char* who) {
    if (!h) return sv_fail(SIEVE_ERR_INVALID, std::string(who) + ": null handle");
    if (!h->ml_ready) return sv_fail(SIEVE_ERR_STATE, std::string(who) + ": call sieve_ml_trace first (and again after any change)");
    return SIEVE_OK;
}

// The max-sum pass over the sites [site_begin, site_begin + n_sites) of the shard into the (chunk-sized) pools.
static int sv_ml_trace_range(sieve_core* h, long long site_begin, int n_sites) {
    const size_t n = h->n, S = (size_t)n_sites, K = h->K, lanes = S * K;
    SvMlLeafArgs la;
    la.counts = h->d_counts;
    la.dmA = h->d_dmA;
    la.dmB = h->d_dmB;
    la.cellL = h->d_mlcellL;
    la.size_factors = h->d_sf;
    la.leafL = h->d_mlL;
    la.ado = h->d_mlado;
    la.n_cells = (int)n;
    la.S = (int)S;
    la.counts_stride = h->S;
    la.site_begin = site_begin;
    la.C = h->C;
    // the trace belongs to the CURRENT leaf buffer: bring the tables to the parameters that buffer was computed with (after
    // a rejected proposal they still hold the rejected values; parameters set since, without a leaf update, do not count)
    {
        const sieve_leaf_params_t pending = h->lp;
        h->lp = h->leaf_lp[h->cur_leaf];
        const int rr = sv_refresh_leaf_tables(h, &la.al, &la.cp);
        h->lp = pending;
        SV_TRY(rr);
    }
    if (h->private_cov) {
        if (!h->have_tv) return sv_fail(SIEVE_ERR_STATE, "sieve_ml_trace: sieve_set_private_seq_cov has not been called");
        SvMlLeafPrivArgs pa;
        pa.counts = h->d_counts;
        pa.dmA = h->d_dmA;
        pa.dmB = h->d_dmB;
        pa.size_factors = h->d_sf;
        pa.tv = h->d_tv;
        pa.leafL = h->d_mlL;
        pa.ado = h->d_mlado;
        pa.n_cells = (int)n;
        pa.S = (int)S;
        pa.C = h->C;
        pa.K = (int)K;
        pa.al = la.al;
        pa.theta = la.cp.theta;
        pa.single_ado = la.cp.single_ado;
        k_leaf_ml_private<<<dim3((unsigned)((S + 127) / 128), (unsigned)n), 128, 0, h->stream>>>(pa);
        sv_count(h);
    } else {
        if (!h->ml_cell_tables_ok) {
            k_build_cell_log_tables<<<dim3((unsigned)((h->C + 127) / 128), (unsigned)n), 128, 0, h->stream>>>(h->d_mlcellL, h->d_sf,
                                                                                                          (int)n, h->C, la.cp);
            sv_count(h);
            h->ml_cell_tables_ok = true;  // per sieve_ml_trace / chunk loop: reset by the callers
        }
        k_leaf_ml<<<dim3((unsigned)((S + 127) / 128), (unsigned)n), 128, 0, h->stream>>>(la);
        sv_count(h);
    }

    SvMlArgs a;
    a.ops = h->d_mlops;
    a.n_ops = h->ml_n_ops;
    a.leafL = h->d_mlL;
    a.ado = h->d_mlado;
    a.logm = h->d_mllogm;
    a.ml = h->d_ml;
    a.back = h->d_mlback;
    a.coll = h->d_mlcoll;
    a.ado_sel = h->d_mladosel;
    a.ambiguous = h->d_mlamb;
    a.S = (int)S;
    a.K = (int)K;
    a.n_leaves = (int)n;
    a.root_genotype = h->root_genotype;
    a.leaf_k = h->private_cov ? 1 : 0;
    a.rawm = h->d_mlrawm;
    a.linear = h->log_mode ? 0 : 1;
    const unsigned blocks = (unsigned)((lanes + 127) / 128);
    k_ml_up<<<blocks, 128, 0, h->stream>>>(a);
    sv_count(h);
    k_ml_down<<<blocks, 128, 0, h->stream>>>(a);
    sv_count(h);
    // the root's sum-product partial (always resident) decides the ML category
    const size_t rslot = (size_t)h->slot_of[(size_t)h->cur_part[h->root] * h->n_nodes + h->root];
    k_ml_categories<<<(unsigned)((S + 255) / 256), 256, 0, h->stream>>>(
        h->d_part + (rslot * (size_t)h->S + (size_t)site_begin) * SV_KMAX * 4, (int)S, (int)K, h->root_genotype, h->d_mlcats);
    sv_count(h);
    SV_CUDA(cudaGetLastError());
    h->ml_site_begin = site_begin;
    h->ml_n_sites = (int)S;
    return SIEVE_OK;
}

// Pools of the max-sum pass.  A dense core traces the whole shard at once (ml_cap == S: every getter works on the result);
// a sparse core has the device memory for a chunk of sites only -- ~190 B per (cell, site) of max-sum partials, back-pointers
// and genotype sets against the 32 B of a leaf -- so sieve_ml_call walks the shard chunk by chunk and keeps the calls.
static int sv_ml_prepare(sieve_core* h) {
    if (h->ephemeral)
        return sv_fail(SIEVE_ERR_STATE, "sieve_ml_trace: needs resident leaves and a resident root (not SIEVE_FLAG_EPHEMERAL)");
    if (!h->leaves_ready) return sv_fail(SIEVE_ERR_STATE, "sieve_ml_trace: leaves not computed yet");
    SV_CUDA(cudaSetDevice(h->device));
    SV_TRY(sv_flush(h));
    const size_t n = h->n, S = h->S, K = h->K, N = h->n_nodes;
    std::vector<SvMlOp> ops;
    SV_TRY(sv_ml_build_ops(h, ops));
    SV_REQUIRE(!ops.empty() && ops.back().dst == h->root - (int)n, "sieve_ml_trace: the root is not the last node of the tree");
    // the root's sum-product partial decides the ML category; make sure it is materialised
    if (!h->valid[h->root]) {
        h->demand_list.push_back(h->root);
        SV_TRY(sv_flush(h));
    }
    if (!h->ml_alloc) {
        size_t cap = S;
        if (h->sparse) {
            const size_t pl = h->private_cov ? K : 1;
            const size_t per_site = n * (pl * 34 + K * (32 + 4 + 1)) + N * K + K + 1;
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            cap = (size_t)(0.7 * (double)free_b / (double)per_site);
            if (const char* e = getenv("SIEVE_B200_ML_CHUNK")) cap = (size_t)std::max(1, atoi(e));
            cap = std::min(S, std::max<size_t>(cap / 128 * 128, 128));
        }
        const size_t lanes = cap * K;
        SV_TRY(sv_alloc(h, &h->d_mlL, n * cap * 4 * (h->private_cov ? K : 1)));
        SV_TRY(sv_alloc(h, &h->d_mlado, n * cap * (h->private_cov ? K : 1)));
        SV_TRY(sv_alloc(h, &h->d_ml, n * lanes * 4));
        SV_TRY(sv_alloc(h, &h->d_mlback, n * lanes));
        SV_TRY(sv_alloc(h, &h->d_mlcoll, N * lanes));
        SV_TRY(sv_alloc(h, &h->d_mladosel, n * lanes));
        SV_TRY(sv_alloc(h, &h->d_mlamb, lanes));
        SV_TRY(sv_alloc(h, &h->d_mlcats, cap));
        SV_TRY(sv_alloc(h, &h->d_mllogm, N * K * 16));
        if (!h->log_mode) SV_TRY(sv_alloc(h, &h->d_mlrawm, N * K * 16));
        SV_TRY(sv_alloc(h, &h->d_mlops, n));
        if (!h->private_cov) SV_TRY(sv_alloc(h, &h->d_mlcellL, n * (size_t)h->C * 4));
        h->ml_cap = (int)cap;
        h->ml_alloc = true;
    }
    // log P on the host, in double precision with the reference's clamp (ScsBeerLikelihoodCore.java:3100-3101)
    std::vector<double> mats((size_t)2 * N * SV_KMAX * 16), logm(N * K * 16);
    SV_CUDA(cudaMemcpyAsync(mats.data(), h->d_mats, mats.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t v = 0; v < N; v++)
        for (size_t k = 0; k < K; k++)
            for (int e = 0; e < 16; e++) {
                const double P = mats[((size_t)h->cur_mat[v] * N + v) * SV_KMAX * 16 + k * 16 + e];
                logm[(v * K + k) * 16 + e] = std::log(P > 0.0 ? P : SV_MIN_VALUE);
            }
    SV_CUDA(cudaMemcpyAsync(h->d_mllogm, logm.data(), logm.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    std::vector<double> rawm;
    if (!h->log_mode) {  // the linear-space twin multiplies a leaf child's likelihoods with P itself (maxsum.cuh)
        rawm.resize(N * K * 16);
        for (size_t v = 0; v < N; v++)
            for (size_t k = 0; k < K; k++)
                memcpy(&rawm[(v * K + k) * 16], &mats[((size_t)h->cur_mat[v] * N + v) * SV_KMAX * 16 + k * 16], 16 * sizeof(double));
        SV_CUDA(cudaMemcpyAsync(h->d_mlrawm, rawm.data(), rawm.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    SV_CUDA(cudaMemcpyAsync(h->d_mlops, ops.data(), ops.size() * sizeof(SvMlOp), cudaMemcpyHostToDevice, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));  // logm / ops are locals
    h->ml_n_ops = (int)ops.size();
    h->ml_cell_tables_ok = false;
    return SIEVE_OK;
}

extern "C" int sieve_ml_trace(sieve_core_t* h) {
    SV_REQUIRE(h, "null handle");
    SV_TRY(sv_ml_prepare(h));
    if (h->ml_cap >= h->S) {
        SV_TRY(sv_ml_trace_range(h, 0, h->S));
        SV_CUDA(cudaStreamSynchronize(h->stream));
    } else
        h->ml_n_sites = 0;  // a chunked core traces inside sieve_ml_call / sieve_ml_get_categories, chunk by chunk
    h->ml_ready = true;
    return SIEVE_OK;
}

// getters that need the whole shard's trace in memory
static int sv_ml_require_whole(sieve_core* h, const char* who) {
    SV_TRY(sv_ml_require(h, who));
    if (h->ml_cap < h->S)
        return sv_fail(SIEVE_ERR_STATE, std::string(who) + ": this core traces in site chunks (sparse residency); per-node and per-site "
                                            "read-backs need a dense core (e.g. one over the tied sites only)");
    return SIEVE_OK;
}

extern "C" int sieve_ml_call(sieve_core_t* h, int8_t* genotypes, int8_t* ado, uint8_t* category, uint8_t* ambiguous) {
    SV_TRY(sv_ml_require(h, "sieve_ml_call"));
    SV_REQUIRE(genotypes && ado && category && ambiguous, "sieve_ml_call: null output");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t n = h->n, S = h->S, N = h->n_nodes, cap = (size_t)h->ml_cap;
    signed char *d_g = nullptr, *d_a = nullptr;
    unsigned char *d_c = nullptr, *d_m = nullptr;
    cudaError_t e = cudaMalloc(&d_g, N * cap);
    if (e == cudaSuccess) e = cudaMalloc(&d_a, n * cap);
    if (e == cudaSuccess) e = cudaMalloc(&d_c, cap);
    if (e == cudaSuccess) e = cudaMalloc(&d_m, cap);
    int rc = SIEVE_OK;
    for (size_t s0 = 0; s0 < S && e == cudaSuccess && rc == SIEVE_OK; s0 += cap) {
        const size_t sc = std::min(cap, S - s0);
        if (cap < S) rc = sv_ml_trace_range(h, (long long)s0, (int)sc);  // a dense core was traced by sieve_ml_trace
        if (rc != SIEVE_OK) break;
        k_ml_call<<<dim3((unsigned)((sc + 255) / 256), (unsigned)N), 256, 0, h->stream>>>(
            h->d_mlcoll, h->d_mladosel, h->d_mlcats, h->d_mlamb, (int)sc, h->K, (int)N, (int)n, d_g, d_a, d_c, d_m);
        sv_count(h);
        e = cudaGetLastError();
        // device [node][sc] -> host [node][S] at column s0
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(genotypes + s0, S, d_g, sc, sc, N, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(ado + s0, S, d_a, sc, sc, n, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(category + s0, d_c, sc, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ambiguous + s0, d_m, sc, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_g);
    cudaFree(d_a);
    cudaFree(d_c);
    cudaFree(d_m);
    SV_TRY(rc);
    SV_CUDA(e);
    return SIEVE_OK;
}

extern "C" int sieve_ml_get_categories(sieve_core_t* h, uint8_t* out) {
    SV_TRY(sv_ml_require(h, "sieve_ml_get_categories"));
    SV_REQUIRE(out, "null output");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t S = h->S, cap = (size_t)h->ml_cap;
    if (cap >= S) {
        SV_CUDA(cudaMemcpyAsync(out, h->d_mlcats, S, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
        return SIEVE_OK;
    }
    // chunked core: the categories come from the root's sum-product partial alone
    const size_t rslot = (size_t)h->slot_of[(size_t)h->cur_part[h->root] * h->n_nodes + h->root];
    for (size_t s0 = 0; s0 < S; s0 += cap) {
        const size_t sc = std::min(cap, S - s0);
        k_ml_categories<<<(unsigned)((sc + 255) / 256), 256, 0, h->stream>>>(h->d_part + (rslot * S + s0) * SV_KMAX * 4, (int)sc, h->K,
                                                                           h->root_genotype, h->d_mlcats);
        sv_count(h);
        SV_CUDA(cudaGetLastError());
        SV_CUDA(cudaMemcpyAsync(out + s0, h->d_mlcats, sc, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
    }
    return SIEVE_OK;
}

// Reference layouts: ml_partials [K][S][4] (MLPartials[..][node - nrOfLeafNodes]), back [K][S][4] (one byte per
// MLGenotypesNodes[..][node][(k, s, parent genotype)]: low nibble = list of child 0, high nibble = list of child 1; for a
// leaf the byte is the list of arg-max dropout components), collection [K][S] (MLGenotypesCollection).
extern "C" int sieve_ml_get_node(sieve_core_t* h, int32_t node, double* ml_partials, uint8_t* back, uint8_t* collection) {
    SV_TRY(sv_ml_require_whole(h, "sieve_ml_get_node"));
    SV_REQUIRE(sv_node_ok(h, node), "sieve_ml_get_node: bad node");
    SV_CUDA(cudaSetDevice(h->device));
    const size_t n = h->n, S = h->S, K = h->K, lanes = S * K;
    if (collection) {
        std::vector<unsigned char> c(lanes);
        SV_CUDA(cudaMemcpyAsync(c.data(), h->d_mlcoll + (size_t)node * lanes, lanes, cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
        for (size_t s = 0; s < S; s++)
            for (size_t k = 0; k < K; k++) collection[k * S + s] = c[s * K + k];
    }
    if ((size_t)node < n) {
        const size_t per = h->private_cov ? K : 1;  // per-category leaves under the private coverage model: [S][K]
        if (ml_partials) {
            std::vector<double> v(S * per * 4);
            SV_CUDA(cudaMemcpyAsync(v.data(), h->d_mlL + (size_t)node * S * per * 4, v.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            SV_CUDA(cudaStreamSynchronize(h->stream));
            for (size_t k = 0; k < K; k++)
                for (size_t s = 0; s < S; s++)
                    memcpy(ml_partials + (k * S + s) * 4, v.data() + (per == 1 ? s : s * K + k) * 4, 4 * sizeof(double));
        }
        if (back) {
            std::vector<unsigned short> v(S * per);
            SV_CUDA(cudaMemcpyAsync(v.data(), h->d_mlado + (size_t)node * S * per, v.size() * sizeof(unsigned short), cudaMemcpyDeviceToHost, h->stream));
            SV_CUDA(cudaStreamSynchronize(h->stream));
            for (size_t k = 0; k < K; k++)
                for (size_t s = 0; s < S; s++)
                    for (int g = 0; g < 4; g++) back[(k * S + s) * 4 + g] = (uint8_t)((v[per == 1 ? s : s * K + k] >> (4 * g)) & 15u);
        }
        return SIEVE_OK;
    }
    const size_t idx = (size_t)node - n;
    if (ml_partials) {
        std::vector<double> v(lanes * 4);
        SV_CUDA(cudaMemcpyAsync(v.data(), h->d_ml + idx * lanes * 4, lanes * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
        for (size_t s = 0; s < S; s++)
            for (size_t k = 0; k < K; k++) memcpy(ml_partials + (k * S + s) * 4, v.data() + (s * K + k) * 4, 4 * sizeof(double));
    }
    if (back) {
        std::vector<unsigned> v(lanes);
        SV_CUDA(cudaMemcpyAsync(v.data(), h->d_mlback + idx * lanes, lanes * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        SV_CUDA(cudaStreamSynchronize(h->stream));
        for (size_t s = 0; s < S; s++)
            for (size_t k = 0; k < K; k++) memcpy(back + (k * S + s) * 4, &v[s * K + k], 4);
    }
    return SIEVE_OK;
}

// One site, every node: back-pointer bytes [node][K][4] and sum-product log partials [node][K][4] -- what a host needs
// to enumerate and rank the tied combinations of an ambiguous site (ScsTreeLikelihood.java:2083-2160).
extern "C" int sieve_ml_get_site(sieve_core_t* h, int32_t site, uint8_t* back, double* log_partials) {
    SV_TRY(sv_ml_require_whole(h, "sieve_ml_get_site"));
    SV_REQUIRE(site >= 0 && site < h->S && back && log_partials, "sieve_ml_get_site: bad argument");
    SV_CUDA(cudaSetDevice(h->device));
    const int n = h->n, N = h->n_nodes, K = h->K;
    // every internal partial has to be in memory: recompute the transient ones (once; they stay valid)
    bool missing = false;
    for (int v = n; v < N; v++)
        if (!h->valid[v]) {
            h->demand_list.push_back(v);
            missing = true;
        }
    if (missing) {
        const bool keep = h->ml_ready;
        SV_TRY(sv_flush(h));
        h->ml_ready = keep;  // recomputing a partial into its buffer changes no value
    }
    std::vector<unsigned> rows(N, 0u);
    for (int v = n; v < N; v++) rows[v] = (unsigned)(((size_t)h->cur_part[v] * h->n_internal + (v - n)) * h->S);
    unsigned* d_rows = nullptr;
    unsigned char* d_b = nullptr;
    double* d_p = nullptr;
    const size_t cnt = (size_t)N * K * 4;
    cudaError_t e = cudaMalloc(&d_rows, N * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&d_b, cnt);
    if (e == cudaSuccess) e = cudaMalloc(&d_p, cnt * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rows, rows.data(), N * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        k_ml_site_gather<<<(N * K + 127) / 128, 128, 0, h->stream>>>(h->d_mlback, h->d_mlado, h->d_mlL, h->d_part, h->d_scale,
                                                                   d_rows, h->S, K, n, N, site, h->private_cov ? 1 : 0, d_b, d_p);
        sv_count(h);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(back, d_b, cnt, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(log_partials, d_p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_rows);
    cudaFree(d_b);
    cudaFree(d_p);
    SV_CUDA(e);
    return SIEVE_OK;
}

// ---- PrivateAllelicSeqCovModel: accept-time re-estimation (leaf_private.cuh, k_private_update) ----------------------
extern "C" int sieve_private_update_from_ml(sieve_core_t* h, int32_t zero_cov_mode, int32_t* out_changed, uint8_t* out_tied) {
    SV_REQUIRE(h && out_changed, "sieve_private_update_from_ml: null argument");
    SV_REQUIRE(zero_cov_mode == 0 || zero_cov_mode == 1, "Illegal value for zeroCovMode. Only 0 or 1 is supported.");
    if (!h->private_cov || !h->have_tv)
        return sv_fail(SIEVE_ERR_STATE, "sieve_private_update_from_ml: needs a SIEVE_FLAG_PRIVATE_SEQ_COV core with its (t, v) set");
    SV_TRY(sieve_ml_trace(h));  // ScsTreeLikelihood.postProcess step 1 (:2454-2463)
    const size_t S = h->S, K = h->K;
    int* d_changed = nullptr;
    unsigned char* d_tied = nullptr;
    cudaError_t e = cudaMalloc(&d_changed, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_tied, K * S);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_changed, 0, sizeof(int), h->stream);
    int changed = 0;
    if (e == cudaSuccess) {
        SvPrivUpdArgs a;
        a.counts = h->d_counts;
        a.size_factors = h->d_sf;
        a.ado_sel = h->d_mladosel;
        a.ambiguous = h->d_mlamb;
        a.tv = h->d_tv;
        a.tied = d_tied;
        a.changed = d_changed;
        a.n_cells = h->n;
        a.S = (int)S;
        a.K = (int)K;
        a.zero_cov_mode = zero_cov_mode;
        a.single_ado = h->single_ado ? 1 : 0;
        k_private_update<<<(unsigned)((S * K + 127) / 128), 128, 0, h->stream>>>(a);
        sv_count(h);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && out_tied) e = cudaMemcpyAsync(out_tied, d_tied, K * S, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_changed);
    cudaFree(d_tied);
    SV_CUDA(e);
    *out_changed = changed;
    if (changed) h->ml_ready = false;  // the trace belongs to the old (t, v)
    return SIEVE_OK;
}

// ScsTreeLikelihood.postProcess step 3 (likelihood/ScsTreeLikelihood.java:2493-2498): "traverse the entire tree for changed
// patterns" -- the changedPatterns overloads of the core (ScsBeerLikelihoodCore.java:584-820, 1070-1266, 1516-1757) restated
// as ONE site list for the device: the sites at which any category's (t, v) differs from the pair the current leaves were
// computed with (after sieve_private_update_from_ml and / or sieve_set_private_seq_cov).  Their leaves, the whole tree
// above them, their per-site log-likelihoods and column sums are re-evaluated IN PLACE in the current buffers -- as in
// the reference no buffer is flipped: every other site keeps its value, which is what a whole-shard pass would
// reproduce bit for bit.  The accepted state then becomes the stored state (sieve_store), since buffers the stored state
// shares with the current one have changed.  ops: the whole tree in post order, as for sieve_update_nodes.
extern "C" int sieve_private_retraverse_changed(sieve_core_t* h, int32_t n_ops, const int32_t* ops, int32_t* out_sites) {
    SV_REQUIRE(h && out_sites && n_ops > 0 && ops, "sieve_private_retraverse_changed: bad argument");
    if (!h->private_cov || !h->have_tv)
        return sv_fail(SIEVE_ERR_STATE, "sieve_private_retraverse_changed: needs a SIEVE_FLAG_PRIVATE_SEQ_COV core with its (t, v) set");
    if (h->sparse || h->ephemeral)
        return sv_fail(SIEVE_ERR_STATE, "sieve_private_retraverse_changed: needs a dense resident core");
    if (!h->leaves_ready || !h->tv_leaf_ok)
        return sv_fail(SIEVE_ERR_STATE, "sieve_private_retraverse_changed: the leaves have not been computed yet");
    if (memcmp(&h->lp, &h->leaf_lp[h->cur_leaf], sizeof(h->lp)) != 0)
        return sv_fail(SIEVE_ERR_STATE,
                       "sieve_private_retraverse_changed: leaf parameters were set since the leaves were computed (the listed sites "
                       "would be evaluated under other parameters than the rest): call sieve_update_leaves first");
    for (int i = 0; i < n_ops; i++)
        SV_REQUIRE(ops[3 * i] >= h->n && ops[3 * i] < h->n_nodes, "sieve_private_retraverse_changed: parent must be an internal node");
    SV_CUDA(cudaSetDevice(h->device));
    SV_TRY(sv_flush(h));  // nothing queued earlier may be mixed into the partial pass
    const int S = h->S;
    unsigned char* d_flag = nullptr;
    int *d_list = nullptr, *d_n = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int n_list = 0;
    thrust::counting_iterator<int> sites(0);
    cudaError_t e = cudaMalloc(&d_flag, S);
    if (e == cudaSuccess) e = cudaMalloc(&d_list, sizeof(int) * S);
    if (e == cudaSuccess) e = cudaMalloc(&d_n, sizeof(int));
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, sites, d_flag, d_list, d_n, S, h->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) {
        if (h->site_out_stale)
            // the per-site outputs belong to a rejected proposal (sieve_restore since the last evaluation: not a sequence the
            // reference produces, accept follows an accepted evaluation): every site is re-evaluated
            e = cudaMemsetAsync(d_flag, 1, S, h->stream);
        else {
            k_tv_changed<<<(S + 255) / 256, 256, 0, h->stream>>>(h->d_tv, h->d_tv_leaf, S, h->K, d_flag);
            sv_count(h);
        }
        if (e == cudaSuccess) {
            e = cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, sites, d_flag, d_list, d_n, S, h->stream);  // ascending site order
            sv_count(h);
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n_list, d_n, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    int r = SIEVE_OK;
    if (e == cudaSuccess && n_list > 0) {
        // leaves of the listed sites into the CURRENT leaf buffer (its other sites hold exactly what a full pass would write)
        h->ml_ready = false;
        h->leaf_ver[h->cur_leaf] = ++h->ver_counter;
        if (h->single_leaf) h->leaf_changed_since_store = true;
        const bool algebraic = h->with_const && !h->const_twin;
        r = sv_launch_leaf_private(h, h->d_leafX + (size_t)h->cur_leaf * h->n * h->S * SV_KMAX * 4, nullptr, h->n, h->d_sumS,
                                   h->d_sum0K, h->d_lsum + (size_t)h->cur_leaf * h->S,
                                   algebraic ? h->d_lambdaK + (size_t)h->cur_leaf * SV_KMAX * h->S : nullptr, h->S, d_list, n_list);
        // every internal node, in place (no flip), over the listed sites
        for (int i = 0; r == SIEVE_OK && i < n_ops; i++) r = sv_record_children(h, ops[3 * i + 1], ops[3 * i + 2], ops[3 * i]);
        if (r == SIEVE_OK) {
            for (int i = 0; i < n_ops; i++) h->dirty_list.push_back(ops[3 * i]);
            h->partial_list = d_list;
            h->partial_n = n_list;
            r = sv_flush(h);
            h->partial_list = nullptr;
            h->partial_n = 0;
            if (r == SIEVE_OK && n_list == S) h->site_out_stale = false;
        }
        if (r == SIEVE_OK) e = cudaStreamSynchronize(h->stream);  // the list is freed below
        if (r == SIEVE_OK) r = sieve_store(h);
    }
    cudaFree(d_flag);
    cudaFree(d_list);
    cudaFree(d_n);
    cudaFree(d_tmp);
    if (r != SIEVE_OK) return r;
    SV_CUDA(e);
    *out_sites = n_list;
    return SIEVE_OK;
}
