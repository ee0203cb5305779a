// Native index builders of the NeoX / Megatron data pipeline (host C++17, pybind11).
//
// Same observable behaviour as the reference's `megatron_dataset/helpers.cpp` (golden outputs in
// tests/test_neox_data.py):
//   build_sample_idx_int32/int64   GPT sample index: (position in doc_idx, token offset) per sample boundary   (helpers.cpp:91-259)
//   build_blending_indices         greedy largest-deficit interleaving of weighted datasets                     (helpers.cpp:34-89)
//   build_mapping                  BERT-style sentence-span samples (seeded target lengths + Fisher-Yates)      (helpers.cpp:261-511)
//   build_blocks_mapping           REALM-style blocks (title-aware target length, per-epoch block ids)          (helpers.cpp:513-745)
// Written as single-pass walkers over growable buffers; ownership passes to numpy through a capsule.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

namespace {

constexpr int32_t kLongSentence = 512;  // documents containing a longer sentence are skipped

// Hand a std::vector over to numpy without copying.
template <typename T>
py::array_t<T> adopt(std::vector<T>&& data, std::vector<py::ssize_t> shape) {
  auto* heap = new std::vector<T>(std::move(data));
  py::capsule owner(heap, [](void* p) { delete reinterpret_cast<std::vector<T>*>(p); });
  std::vector<py::ssize_t> strides(shape.size());
  py::ssize_t s = sizeof(T);
  for (int i = int(shape.size()) - 1; i >= 0; --i) {
    strides[i] = s;
    s *= shape[i];
  }
  return py::array_t<T>(shape, strides, heap->data(), owner);
}

// ------------------------------------------------------------------------------------------- GPT sample index
// Documents are concatenated in `doc_idx` order; every sample covers seq_length + 1 tokens and consecutive samples
// overlap by one token.  Entry i is where sample i starts: (index into doc_idx, offset inside that document).
template <typename Idx>
py::array_t<Idx> sample_index(const py::array_t<int32_t>& sizes_, const py::array_t<int32_t>& doc_idx_, int32_t seq_length,
                              int32_t num_epochs, int64_t tokens_per_epoch) {
  if (seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) throw std::invalid_argument("bad sample-index arguments");
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = doc_idx_.unchecked<1>();
  const int64_t n_samples = (int64_t(num_epochs) * tokens_per_epoch - 1) / seq_length;
  // (python-level print: iostream from an extension loaded next to libtorch's bundled libstdc++ is fragile)
  py::print("    using:");
  py::print("     number of documents:      ", doc_idx_.shape(0) / num_epochs);
  py::print("     number of epochs:         ", num_epochs);
  py::print("     sequence length:          ", seq_length);
  py::print("     total number of samples:  ", n_samples);
  std::vector<Idx> out;
  out.reserve(size_t(2 * (n_samples + 1)));
  int64_t cursor = 0;   // position in doc_idx
  int32_t offset = 0;   // first unread token of that document
  out.push_back(Idx(cursor));
  out.push_back(Idx(offset));
  for (int64_t s = 0; s < n_samples; ++s) {
    int32_t need = seq_length + 1;
    while (need > 0) {
      const int32_t avail = sizes[doc_idx[cursor]] - offset;
      if (avail >= need) {
        offset += need - 1;  // the last token of this sample is re-read as the first of the next one
        need = 0;
      } else {
        need -= avail;
        ++cursor;
        offset = 0;
      }
    }
    out.push_back(Idx(cursor));
    out.push_back(Idx(offset));
  }
  return adopt(std::move(out), {py::ssize_t(n_samples + 1), 2});
}

// ------------------------------------------------------------------------------------------- dataset blending
// At step t pick the dataset whose achieved count lags its target weight*t the most.
void blending_indices(py::array_t<uint8_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index,
                      const py::array_t<double>& weights, int32_t num_datasets, int64_t size, bool verbose) {
  auto which = dataset_index.mutable_unchecked<1>();
  auto where = dataset_sample_index.mutable_unchecked<1>();
  auto w = weights.unchecked<1>();
  std::vector<int64_t> taken(size_t(num_datasets), 0);
  for (int64_t t = 0; t < size; ++t) {
    const double denom = std::max(double(t), 1.0);
    int64_t best = 0;
    double best_gap = w[0] * denom - double(taken[0]);
    for (int64_t d = 1; d < num_datasets; ++d) {
      const double gap = w[d] * denom - double(taken[size_t(d)]);
      if (gap > best_gap) {
        best_gap = gap;
        best = d;
      }
    }
    which[t] = uint8_t(best);
    where[t] = taken[size_t(best)]++;
  }
  if (verbose) {
    py::print(" > sample ratios:");
    for (int64_t d = 0; d < num_datasets; ++d)
      py::print("   dataset", d, ", input:", w[d], ", achieved:", double(taken[size_t(d)]) / double(size));
  }
}

// ------------------------------------------------------------------------------------------- sentence-span maps
template <typename Idx, int Width>
void fisher_yates(std::vector<Idx>& rows, int32_t seed) {
  std::mt19937_64 gen{static_cast<uint64_t>(seed + 1)};
  const int64_t n = int64_t(rows.size()) / Width;
  for (int64_t i = n - 1; i > 0; --i) {
    const int64_t j = int64_t(gen() % uint64_t(i + 1));
    for (int c = 0; c < Width; ++c) std::swap(rows[size_t(Width * i + c)], rows[size_t(Width * j + c)]);
  }
}

inline bool has_long_sentence(const py::detail::unchecked_reference<int32_t, 1>& sizes, int64_t first, int64_t last) {
  for (int64_t s = first; s < last; ++s)
    if (sizes[s] > kLongSentence) return true;
  return false;
}

template <typename Idx>
py::array mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, int32_t num_epochs, uint64_t max_num_samples,
                  int32_t max_seq_length, double short_seq_prob, int32_t seed, bool verbose) {
  if (num_epochs <= 0 || max_seq_length <= 1 || !(short_seq_prob > 0.0 && short_seq_prob <= 1.0) || seed <= 0)
    throw std::invalid_argument("bad build_mapping arguments");
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  const int32_t short_ratio = int32_t(std::lround(1.0 / short_seq_prob));
  std::mt19937 gen{static_cast<uint32_t>(seed)};
  auto draw_target = [&]() -> int32_t {
    const auto x = gen();
    return (x % uint32_t(short_ratio)) == 0 ? int32_t(2 + x % uint32_t(max_seq_length - 1)) : max_seq_length;
  };
  std::vector<Idx> rows;
  uint64_t n_rows = 0;
  const int64_t n_docs = docs_.shape(0) - 1;
  for (int32_t epoch = 0; epoch < num_epochs && n_rows < max_num_samples; ++epoch) {
    for (int64_t d = 0; d < n_docs; ++d) {
      const int64_t first = docs[d], last = docs[d + 1];
      int64_t remaining = last - first;
      if (remaining <= 1 || has_long_sentence(sizes, first, last)) continue;
      int64_t span_start = first;
      int32_t span_tokens = 0, span_sents = 0;
      int32_t target = draw_target();
      for (int64_t s = first; s < last; ++s) {
        span_tokens += sizes[s];
        ++span_sents;
        --remaining;
        const bool full = span_tokens >= target && remaining > 1 && span_sents > 1;
        if (full || remaining == 0) {
          rows.push_back(Idx(span_start));
          rows.push_back(Idx(s + 1));
          rows.push_back(Idx(target));
          ++n_rows;
          span_start = s + 1;
          target = draw_target();
          span_tokens = 0;
          span_sents = 0;
        }
      }
    }
  }
  if (verbose) py::print("   will create mapping for", n_rows, "samples");
  fisher_yates<Idx, 3>(rows, seed);
  return adopt(std::move(rows), {py::ssize_t(n_rows), 3});
}

template <typename Idx>
py::array blocks_mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, const py::array_t<int32_t>& titles_,
                         int32_t num_epochs, uint64_t max_num_samples, int32_t max_seq_length, int32_t seed, bool verbose,
                         bool use_one_sent_blocks) {
  if (num_epochs <= 0 || max_seq_length <= 1 || seed <= 0) throw std::invalid_argument("bad build_blocks_mapping arguments");
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  auto titles = titles_.unchecked<1>();
  const int64_t min_sents = use_one_sent_blocks ? 1 : 2;
  std::vector<Idx> rows;
  uint64_t n_rows = 0;
  const int64_t n_docs = docs_.shape(0) - 1;
  for (int32_t epoch = 0; epoch < num_epochs && n_rows < max_num_samples; ++epoch) {
    int32_t block_id = 0;  // ids restart every epoch
    for (int64_t d = 0; d < n_docs; ++d) {
      const int64_t first = docs[d], last = docs[d + 1];
      int64_t remaining = last - first;
      if (remaining < min_sents || has_long_sentence(sizes, first, last)) continue;
      const int32_t target = max_seq_length - titles[d];
      int64_t span_start = first;
      int32_t span_tokens = 0;
      int64_t span_sents = 0;
      for (int64_t s = first; s < last; ++s) {
        span_tokens += sizes[s];
        ++span_sents;
        --remaining;
        const bool full = span_tokens >= target && remaining >= min_sents && span_sents >= min_sents;
        if (full || remaining == 0) {
          rows.push_back(Idx(span_start));
          rows.push_back(Idx(s + 1));
          rows.push_back(Idx(d));
          rows.push_back(Idx(block_id++));
          ++n_rows;
          span_start = s + 1;
          span_tokens = 0;
          span_sents = 0;
        }
      }
    }
  }
  if (verbose) py::print("   will create mapping for", n_rows, "samples");
  fisher_yates<Idx, 4>(rows, seed);
  return adopt(std::move(rows), {py::ssize_t(n_rows), 4});
}

py::array build_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes, int num_epochs, uint64_t max_num_samples,
                        int max_seq_length, double short_seq_prob, int seed, bool verbose) {
  if (uint64_t(sizes.size()) > std::numeric_limits<uint32_t>::max())
    return mapping<uint64_t>(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose);
  return mapping<uint32_t>(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose);
}

py::array build_blocks_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes, const py::array_t<int32_t>& titles,
                               int num_epochs, uint64_t max_num_samples, int max_seq_length, int seed, bool verbose,
                               bool use_one_sent_blocks) {
  if (uint64_t(sizes.size()) > std::numeric_limits<uint32_t>::max())
    return blocks_mapping<uint64_t>(docs, sizes, titles, num_epochs, max_num_samples, max_seq_length, seed, verbose, use_one_sent_blocks);
  return blocks_mapping<uint32_t>(docs, sizes, titles, num_epochs, max_num_samples, max_seq_length, seed, verbose, use_one_sent_blocks);
}

}  // namespace

PYBIND11_MODULE(_data_helpers, m) {
  m.doc() = "relora_b200 dataset index builders";
  m.def("build_sample_idx_int32", &sample_index<int32_t>);
  m.def("build_sample_idx_int64", &sample_index<int64_t>);
  m.def("build_blending_indices", &blending_indices);
  m.def("build_mapping", &build_mapping);
  m.def("build_blocks_mapping", &build_blocks_mapping);
}
