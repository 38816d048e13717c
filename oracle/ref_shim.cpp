// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// A small pybind11 module, compiled by oracle/build_ref.sh against the UNMODIFIED reference headers
// under /root/reference/src/cpp, that exposes the reference's header-only / free functions so the
// parity tests and golden-vector generator can call them at unit level:
//   scan_list                 include/list_scanning.h:292-311
//   batched_scan_list         include/list_scanning.h:313-366
//   TopkBuffer                include/list_scanning.h:41-204
//   kmeans                    src/clustering.cpp:13-97
//   kmeans_refine_partitions  src/clustering.cpp:99-182
//   compute_boundary_distances / compute_recall_profile   include/geometry.h:57-113, 345-407
//   MaintenanceCostEstimator (split / delete deltas) + ListScanLatencyEstimator interpolation
//                             src/maintenance_cost_estimator.cpp:131-251, 355-493 -- with a latency table supplied by
//                             the test (the reference's own table is a CPU profile of the host it runs on)
// Nothing here restates reference logic; it only marshals tensors.
#include <torch/extension.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <common.h>
#include <list_scanning.h>
#include <clustering.h>
#include <geometry.h>
#include <index_partition.h>
#include <maintenance_cost_estimator.h>

namespace py = pybind11;

static faiss::MetricType metric_of(const std::string &m) { return str_to_metric_type(m); }

// one query x one list -> (ids, dists), best-first, min(k, n) entries
static std::tuple<Tensor, Tensor> shim_scan_list(Tensor query, Tensor list_vecs, c10::optional<Tensor> list_ids,
                                                  int k, const std::string &metric) {
    query = query.contiguous();
    list_vecs = list_vecs.contiguous();
    bool desc = metric_of(metric) == faiss::METRIC_INNER_PRODUCT;
    TopkBuffer buf(k, desc);
    const int64_t *ids_ptr = nullptr;
    Tensor ids_c;
    if (list_ids.has_value() && list_ids->defined()) {
        ids_c = list_ids->contiguous();
        ids_ptr = ids_c.data_ptr<int64_t>();
    }
    int n = list_vecs.size(0);
    int d = query.size(0);
    scan_list(query.data_ptr<float>(), n ? list_vecs.data_ptr<float>() : nullptr, ids_ptr, n, d, buf,
              metric_of(metric));
    auto dists = buf.get_topk();
    auto ids = buf.get_topk_indices();
    return {torch::tensor(ids, torch::kInt64), torch::tensor(dists, torch::kFloat32)};
}

// G queries x one list -> per-query (ids, dists) padded with -1 / +-inf to k; counts returned too
static std::tuple<Tensor, Tensor, Tensor> shim_batched_scan_list(Tensor queries, Tensor list_vecs,
                                                                  c10::optional<Tensor> list_ids, int k,
                                                                  const std::string &metric) {
    queries = queries.contiguous();
    list_vecs = list_vecs.contiguous();
    bool desc = metric_of(metric) == faiss::METRIC_INNER_PRODUCT;
    int nq = queries.size(0);
    int d = queries.size(1);
    int n = list_vecs.size(0);
    auto bufs = create_buffers(nq, k, desc);
    const int64_t *ids_ptr = nullptr;
    Tensor ids_c;
    if (list_ids.has_value() && list_ids->defined()) {
        ids_c = list_ids->contiguous();
        ids_ptr = ids_c.data_ptr<int64_t>();
    }
    batched_scan_list(queries.data_ptr<float>(), n ? list_vecs.data_ptr<float>() : nullptr, ids_ptr, nq, n, d, bufs,
                      metric_of(metric));
    Tensor out_ids = torch::full({nq, k}, -1, torch::kInt64);
    Tensor out_d = torch::full({nq, k}, desc ? -INFINITY : INFINITY, torch::kFloat32);
    Tensor counts = torch::zeros({nq}, torch::kInt64);
    for (int i = 0; i < nq; i++) {
        auto dd = bufs[i]->get_topk();
        auto ii = bufs[i]->get_topk_indices();
        counts[i] = (int64_t) dd.size();
        for (size_t j = 0; j < dd.size(); j++) {
            out_ids[i][j] = ii[j];
            out_d[i][j] = dd[j];
        }
    }
    return {out_ids, out_d, counts};
}

// TopkBuffer driven by a stream of (dist, id) with an explicit capacity; returns sorted top-k.
static std::tuple<Tensor, Tensor, float> shim_topk_buffer(Tensor dists, Tensor ids, int k, bool descending,
                                                           int capacity) {
    dists = dists.contiguous();
    ids = ids.contiguous();
    TopkBuffer buf(k, descending, capacity);
    auto dp = dists.data_ptr<float>();
    auto ip = ids.data_ptr<int64_t>();
    for (int64_t i = 0; i < dists.numel(); i++) buf.add(dp[i], ip[i]);
    float kth = buf.get_kth_distance();
    auto d = buf.get_topk();
    auto i = buf.get_topk_indices();
    return {torch::tensor(i, torch::kInt64), torch::tensor(d, torch::kFloat32), kth};
}

static std::tuple<Tensor, std::vector<Tensor>, std::vector<Tensor>> shim_kmeans(Tensor x, Tensor ids, int nlist,
                                                                                 const std::string &metric,
                                                                                 int niter) {
    auto c = kmeans(x.contiguous().clone(), ids.contiguous(), nlist, metric_of(metric), niter, false);
    return {c->centroids, c->vectors, c->vector_ids};
}

static std::tuple<Tensor, std::vector<Tensor>, std::vector<Tensor>> shim_kmeans_refine(
    Tensor centroids, std::vector<Tensor> part_vecs, std::vector<Tensor> part_ids, const std::string &metric,
    int iterations) {
    centroids = centroids.contiguous().clone();
    int d = centroids.size(1);
    std::vector<std::shared_ptr<IndexPartition>> parts;
    for (size_t i = 0; i < part_vecs.size(); i++) {
        auto p = std::make_shared<IndexPartition>();
        p->set_code_size(d * sizeof(float));
        Tensor v = part_vecs[i].contiguous();
        Tensor id = part_ids[i].contiguous();
        if (v.size(0) > 0) p->append(v.size(0), id.data_ptr<int64_t>(), (const uint8_t *) v.data_ptr<float>());
        parts.push_back(p);
    }
    auto [new_c, new_parts] = kmeans_refine_partitions(centroids, parts, metric_of(metric), iterations);
    std::vector<Tensor> ov, oi;
    for (auto &p: new_parts) {
        int64_t n = p->num_vectors_;
        Tensor v = torch::empty({n, d}, torch::kFloat32);
        Tensor id = torch::empty({n}, torch::kInt64);
        if (n > 0) {
            std::memcpy(v.data_ptr<float>(), p->codes_, n * d * sizeof(float));
            std::memcpy(id.data_ptr<int64_t>(), p->ids_, n * sizeof(int64_t));
        }
        ov.push_back(v);
        oi.push_back(id);
    }
    return {new_c.clone(), ov, oi};
}

// raw pairwise kernels of the vendored faiss (third_party/faiss/faiss/utils/distances_simd.cpp:188-224)
static Tensor shim_pairwise(Tensor x, Tensor y, const std::string &metric) {
    x = x.contiguous();
    y = y.contiguous();
    int64_t nx = x.size(0), ny = y.size(0), d = x.size(1);
    Tensor out = torch::empty({nx, ny}, torch::kFloat32);
    bool ip = metric_of(metric) == faiss::METRIC_INNER_PRODUCT;
    float *o = out.data_ptr<float>();
    for (int64_t i = 0; i < nx; i++)
        for (int64_t j = 0; j < ny; j++)
            o[i * ny + j] = ip ? faiss::fvec_inner_product(x.data_ptr<float>() + i * d, y.data_ptr<float>() + j * d, d)
                               : faiss::fvec_L2sqr(x.data_ptr<float>() + i * d, y.data_ptr<float>() + j * d, d);
    return out;
}

static std::vector<float> shim_boundary_distances(Tensor query, Tensor centroids, bool euclidean) {
    query = query.contiguous();
    centroids = centroids.contiguous();
    std::vector<float *> ptrs;
    for (int64_t j = 0; j < centroids.size(0); j++) ptrs.push_back(centroids.data_ptr<float>() + j * centroids.size(1));
    return compute_boundary_distances(query, ptrs, euclidean);
}

static std::vector<float> shim_recall_profile(std::vector<float> boundary, float radius, int d, bool use_precomputed,
                                              bool euclidean) {
    return compute_recall_profile(boundary, radius, d, {}, use_precomputed, euclidean);
}

// The reference's cost model with its latency grid values replaced by `table` [n_values x k_values] (public member)
struct ShimCostModel {
    std::shared_ptr<MaintenanceCostEstimator> est;
    ShimCostModel(int d, float alpha, int k, std::vector<std::vector<float>> table) {
        est = std::make_shared<MaintenanceCostEstimator>(d, alpha, k);  // profiles the CPU once (discarded below)
        auto lat = est->get_latency_estimator();
        if (table.size() != lat->n_values_.size() || table[0].size() != lat->k_values_.size())
            throw std::runtime_error("latency table shape does not match the reference's grid");
        lat->scan_latency_model_ = table;
    }
    float latency(int n, int k) const { return est->get_latency_estimator()->estimate_scan_latency(n, k); }
    float split_delta(int size, float hit_rate, int total) const { return est->compute_split_delta(size, hit_rate, total); }
    float delete_delta(int size, float hit_rate, int total, float avg_rate, float avg_size) const {
        return est->compute_delete_delta(size, hit_rate, total, avg_rate, avg_size);
    }
    float delete_delta_w_reassign(int size, float hit_rate, int total, std::vector<int64_t> counts,
                                  std::vector<int64_t> sizes, std::vector<float> rates) const {
        return est->compute_delete_delta_w_reassign(size, hit_rate, total, counts, sizes, rates);
    }
    std::vector<int> n_values() const { return est->get_latency_estimator()->n_values_; }
    std::vector<int> k_values() const { return est->get_latency_estimator()->k_values_; }
};

PYBIND11_MODULE(_shim, m) {
    py::class_<ShimCostModel>(m, "CostModel")
        .def(py::init<int, float, int, std::vector<std::vector<float>>>())
        .def("latency", &ShimCostModel::latency)
        .def("split_delta", &ShimCostModel::split_delta)
        .def("delete_delta", &ShimCostModel::delete_delta)
        .def("delete_delta_w_reassign", &ShimCostModel::delete_delta_w_reassign)
        .def("n_values", &ShimCostModel::n_values)
        .def("k_values", &ShimCostModel::k_values);
    m.def("scan_list", &shim_scan_list, py::arg("query"), py::arg("list_vecs"), py::arg("list_ids"), py::arg("k"),
          py::arg("metric"));
    m.def("batched_scan_list", &shim_batched_scan_list, py::arg("queries"), py::arg("list_vecs"),
          py::arg("list_ids"), py::arg("k"), py::arg("metric"));
    m.def("topk_buffer", &shim_topk_buffer);
    m.def("kmeans", &shim_kmeans);
    m.def("kmeans_refine_partitions", &shim_kmeans_refine);
    m.def("compute_boundary_distances", &shim_boundary_distances);
    m.def("compute_recall_profile", &shim_recall_profile);
    m.def("pairwise", &shim_pairwise);
    m.def("incomplete_beta", &incomplete_beta);
    m.def("incomplete_beta_lookup", &incomplete_beta_lookup);
}
