"""On-disk token stores of the Megatron / GPT-NeoX family: ``<prefix>.bin`` (raw token stream) + ``<prefix>.idx``.

Two index formats are supported, both byte-compatible with the reference
(``megatron_dataset/indexed_dataset.py:133-273`` legacy, ``:348-603`` mmap):

``mmap``   magic ``MMIDIDX\\x00\\x00`` · <Q version=1 · <B dtype code · <Q n_sizes · <Q n_docs ·
           int32 sizes[n] · int64 byte-pointers[n] · int64 doc_idx[n_docs]
``lazy`` / ``cached``  magic ``TNTIDX\\x00\\x00`` · <Q version=1 · <QQ code, element_size · <QQ len, n_sizes · <Q doc_count ·
           int64 dim_offsets[len+1] · int64 data_offsets[len+1] · int64 sizes[n_sizes] · int64 doc_idx[doc_count]

The mmap reader keeps zero-copy numpy views over the index and the token stream; ``get(idx, offset, length)``
returns a slice of one document without touching the rest.
"""
from __future__ import annotations

import os
import shutil
import struct
from functools import lru_cache
from itertools import accumulate
from typing import List, Optional

import numpy as np
import torch

from ...obs import logger

__all__ = [
    "MMapIndexedDataset",
    "MMapIndexedDatasetBuilder",
    "IndexedDataset",
    "IndexedCachedDataset",
    "IndexedDatasetBuilder",
    "make_dataset",
    "make_builder",
    "infer_dataset_impl",
    "dataset_exists",
    "best_fitting_dtype",
    "index_file_path",
    "data_file_path",
    "DTYPES",
]

DTYPES = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float32, 7: np.float64, 8: np.uint16}
_MMAP_MAGIC = b"MMIDIDX\x00\x00"
_LEGACY_MAGIC = b"TNTIDX\x00\x00"


def dtype_code(dtype) -> int:
    for k, v in DTYPES.items():
        if v == dtype:
            return k
    raise ValueError(dtype)


def best_fitting_dtype(vocab_size: Optional[int] = None):
    return np.uint16 if vocab_size is not None and vocab_size < 65500 else np.int32


def index_file_path(prefix: str) -> str:
    return prefix + ".idx"


def data_file_path(prefix: str) -> str:
    return prefix + ".bin"


def infer_dataset_impl(path: str) -> Optional[str]:
    if not dataset_exists(path, "lazy"):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    with open(index_file_path(path), "rb") as f:
        magic = f.read(8)
        if magic == _LEGACY_MAGIC:
            return "cached"
        if magic == _MMAP_MAGIC[:8]:
            return "mmap"
    return None


def dataset_exists(path: str, impl: str) -> bool:
    return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


def make_builder(out_file: str, impl: str, vocab_size: Optional[int] = None):
    if impl == "mmap":
        return MMapIndexedDatasetBuilder(out_file, dtype=best_fitting_dtype(vocab_size))
    return IndexedDatasetBuilder(out_file)


def make_dataset(path: str, impl: str, skip_warmup: bool = False):
    if not dataset_exists(path, impl):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    if impl == "infer":
        impl = infer_dataset_impl(path)
    if impl == "lazy":
        return IndexedDataset(path)
    if impl == "cached":
        return IndexedCachedDataset(path)
    if impl == "mmap":
        return MMapIndexedDataset(path, skip_warmup)
    print(f"Unknown dataset implementation: {impl}")
    return None


def _doc_boundaries(sizes) -> List[int]:
    out = [0]
    for i, s in enumerate(sizes):
        if s == 0:
            out.append(i + 1)
    return out


# ----------------------------------------------------------------------------------------------- legacy format
class IndexedDataset(torch.utils.data.Dataset):
    """Legacy (fairseq-style) indexed dataset read with plain file seeks."""

    def __init__(self, path: str):
        super().__init__()
        self.path = path
        self.data_file = None
        with open(index_file_path(path), "rb") as f:
            if f.read(8) != _LEGACY_MAGIC:
                raise ValueError("Index file doesn't match expected format. Make sure that --dataset-impl is configured properly.")
            (version,) = struct.unpack("<Q", f.read(8))
            assert version == 1
            code, self.element_size = struct.unpack("<QQ", f.read(16))
            self.dtype = DTYPES[code]
            self._len, n_sizes = struct.unpack("<QQ", f.read(16))
            (doc_count,) = struct.unpack("<Q", f.read(8))
            rd = lambda n: np.frombuffer(f.read(8 * n), dtype=np.int64).copy()  # noqa: E731
            self.dim_offsets = rd(self._len + 1)
            self.data_offsets = rd(self._len + 1)
            self.sizes = rd(n_sizes)
            self.doc_idx = rd(doc_count)

    def _open(self):
        if self.data_file is None:
            self.data_file = open(data_file_path(self.path), "rb", buffering=0)

    def __del__(self):
        if getattr(self, "data_file", None):
            self.data_file.close()

    def __len__(self):
        return self._len

    def num_tokens(self, index):
        return self.sizes[index]

    def size(self, index):
        return self.sizes[index]

    @property
    def supports_prefetch(self):
        return False

    def __getitem__(self, idx):
        self._open()
        if isinstance(idx, (int, np.integer)):
            if idx < 0 or idx >= self._len:
                raise IndexError("index out of range")
            shape = self.sizes[self.dim_offsets[idx]: self.dim_offsets[idx + 1]]
            a = np.empty(shape, dtype=self.dtype)
            self.data_file.seek(int(self.data_offsets[idx]) * self.element_size)
            self.data_file.readinto(a)
            return a
        start, stop, step = idx.indices(len(self))
        if step != 1:
            raise ValueError("Slices into indexed_dataset must be contiguous")
        sizes = self.sizes[self.dim_offsets[start]: self.dim_offsets[stop]]
        a = np.empty(int(sum(sizes)), dtype=self.dtype)
        self.data_file.seek(int(self.data_offsets[start]) * self.element_size)
        self.data_file.readinto(a)
        return np.split(a, list(accumulate(sizes))[:-1])


class IndexedCachedDataset(IndexedDataset):
    """Legacy dataset with an explicit prefetch cache (``prefetch(indices)`` then random access)."""

    def __init__(self, path):
        super().__init__(path)
        self.cache = None
        self.cache_index = {}

    @property
    def supports_prefetch(self):
        return True

    def prefetch(self, indices):
        if all(i in self.cache_index for i in indices):
            return
        self._open()
        indices = sorted(set(indices))
        total = sum(int(self.data_offsets[i + 1] - self.data_offsets[i]) for i in indices)
        self.cache = np.empty(total, dtype=self.dtype)
        self.cache_index.clear()
        ptx = 0
        for i in indices:
            self.cache_index[i] = ptx
            size = int(self.data_offsets[i + 1] - self.data_offsets[i])
            self.data_file.seek(int(self.data_offsets[i]) * self.element_size)
            self.data_file.readinto(self.cache[ptx: ptx + size])
            ptx += size
        self.data_file.close()
        self.data_file = None

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            if idx < 0 or idx >= self._len:
                raise IndexError("index out of range")
            shape = self.sizes[self.dim_offsets[idx]: self.dim_offsets[idx + 1]]
            a = np.empty(shape, dtype=self.dtype)
            ptx = self.cache_index[idx]
            np.copyto(a, self.cache[ptx: ptx + a.size].reshape(a.shape))
            return a
        return [self[i] for i in range(*idx.indices(len(self)))]


class IndexedDatasetBuilder:
    element_sizes = {np.uint8: 1, np.int8: 1, np.int16: 2, np.int32: 4, np.int64: 8, np.float32: 4, np.float64: 8}

    def __init__(self, out_file: str, dtype=np.int32):
        self.out_file = open(out_file, "wb")
        self.dtype = dtype
        self.data_offsets, self.dim_offsets, self.sizes, self.doc_idx = [0], [0], [], [0]
        self.element_size = self.element_sizes[self.dtype]

    def add_item(self, arr):
        arr = np.asarray(arr, dtype=self.dtype)
        self.out_file.write(arr.tobytes(order="C"))
        self.data_offsets.append(self.data_offsets[-1] + arr.size)
        self.sizes.extend(arr.shape)
        self.dim_offsets.append(self.dim_offsets[-1] + arr.ndim)

    def end_document(self):
        self.doc_idx.append(len(self.sizes))

    def merge_file_(self, another_file: str):
        other = IndexedDataset(another_file)
        assert other.dtype == self.dtype
        begin = self.data_offsets[-1]
        self.data_offsets.extend(begin + int(o) for o in other.data_offsets[1:])
        self.sizes.extend(int(s) for s in other.sizes)
        begin = self.dim_offsets[-1]
        self.dim_offsets.extend(begin + int(o) for o in other.dim_offsets[1:])
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self.out_file)

    def finalize(self, index_file: str):
        self.out_file.close()
        with open(index_file, "wb") as f:
            f.write(_LEGACY_MAGIC)
            f.write(struct.pack("<Q", 1))
            f.write(struct.pack("<QQ", dtype_code(self.dtype), self.element_size))
            f.write(struct.pack("<QQ", len(self.data_offsets) - 1, len(self.sizes)))
            f.write(struct.pack("<Q", len(self.doc_idx)))
            for arr in (self.dim_offsets, self.data_offsets, self.sizes, self.doc_idx):
                f.write(np.array(arr, dtype=np.int64).tobytes())


# ----------------------------------------------------------------------------------------------- mmap format
class _MMapIndex:
    def __init__(self, path: str, skip_warmup: bool = False):
        with open(path, "rb") as f:
            if f.read(9) != _MMAP_MAGIC:
                raise ValueError("Index file doesn't match expected format. Make sure that --dataset-impl is configured properly.")
            (version,) = struct.unpack("<Q", f.read(8))
            assert version == 1
            (code,) = struct.unpack("<B", f.read(1))
            self.dtype = DTYPES[code]
            self.dtype_size = np.dtype(self.dtype).itemsize
            (self._len,) = struct.unpack("<Q", f.read(8))
            (self._doc_count,) = struct.unpack("<Q", f.read(8))
            offset = f.tell()
        if not skip_warmup:
            logger.info("    warming up index mmap file...")
            _warmup(path)
        self._mmap = np.memmap(path, mode="r", order="C")
        buf = memoryview(self._mmap)
        self.sizes = np.frombuffer(buf, dtype=np.int32, count=self._len, offset=offset)
        self.pointers = np.frombuffer(buf, dtype=np.int64, count=self._len, offset=offset + self.sizes.nbytes)
        self.doc_idx = np.frombuffer(buf, dtype=np.int64, count=self._doc_count, offset=offset + self.sizes.nbytes + self.pointers.nbytes)

    def __len__(self):
        return self._len

    def close(self):
        mm = getattr(self, "_mmap", None)
        if mm is not None and getattr(mm, "_mmap", None) is not None:
            self.sizes = self.pointers = self.doc_idx = None
            try:
                mm._mmap.close()
            except Exception:  # views still alive: let the GC handle it
                pass
        self._mmap = None

    @staticmethod
    def write(path: str, dtype, sizes, doc_idx):
        sizes = np.asarray(sizes, dtype=np.int32)
        itemsize = np.dtype(dtype).itemsize
        pointers = np.zeros(len(sizes), dtype=np.int64)
        if len(sizes) > 1:
            np.cumsum(sizes[:-1].astype(np.int64) * itemsize, out=pointers[1:])
        with open(path, "wb") as f:
            f.write(_MMAP_MAGIC)
            f.write(struct.pack("<Q", 1))
            f.write(struct.pack("<B", dtype_code(dtype)))
            f.write(struct.pack("<Q", len(sizes)))
            f.write(struct.pack("<Q", len(doc_idx)))
            f.write(sizes.tobytes(order="C"))
            f.write(pointers.tobytes(order="C"))
            f.write(np.asarray(doc_idx, dtype=np.int64).tobytes(order="C"))


def _warmup(path: str):
    with open(path, "rb") as stream:
        while stream.read(100 * 1024 * 1024):
            pass


class MMapIndexedDataset(torch.utils.data.Dataset):
    def __init__(self, path: str, skip_warmup: bool = False):
        super().__init__()
        self._path = None
        self._index = None
        self._bin = None
        self._do_init(path, skip_warmup)

    def __getstate__(self):
        return self._path

    def __setstate__(self, state):
        self._do_init(state, skip_warmup=True)

    def _do_init(self, path, skip_warmup):
        self._path = path
        self._index = _MMapIndex(index_file_path(path), skip_warmup)
        if not skip_warmup:
            logger.info("    warming up data mmap file...")
            _warmup(data_file_path(path))
        self._bin = np.memmap(data_file_path(path), mode="r", order="C")
        self._buf = memoryview(self._bin)

    def __len__(self):
        return len(self._index)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            ptr, size = int(self._index.pointers[idx]), int(self._index.sizes[idx])
            return np.frombuffer(self._buf, dtype=self._index.dtype, count=size, offset=ptr)
        start, stop, step = idx.indices(len(self))
        if step != 1:
            raise ValueError("Slices into indexed_dataset must be contiguous")
        ptr = int(self._index.pointers[start])
        sizes = self._index.sizes[idx]
        a = np.frombuffer(self._buf, dtype=self._index.dtype, count=int(sizes.sum()), offset=ptr)
        return np.split(a, list(accumulate(sizes))[:-1])

    def get(self, idx, offset: int = 0, length: Optional[int] = None):
        """``length`` tokens of document ``idx`` starting at ``offset`` (zero-copy)."""
        ptr, size = int(self._index.pointers[idx]), int(self._index.sizes[idx])
        if length is None:
            length = size - offset
        ptr += offset * np.dtype(self._index.dtype).itemsize
        return np.frombuffer(self._buf, dtype=self._index.dtype, count=length, offset=ptr)

    @property
    def sizes(self):
        return self._index.sizes

    @property
    def doc_idx(self):
        return self._index.doc_idx

    def get_doc_idx(self):
        return self._index.doc_idx

    def set_doc_idx(self, doc_idx_):
        self._index.doc_idx = doc_idx_

    @property
    def supports_prefetch(self):
        return False

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


class MMapIndexedDatasetBuilder:
    def __init__(self, out_file: str, dtype=np.int64):
        self._data_file = open(out_file, "wb")
        self._dtype = dtype
        self._sizes: List[int] = []
        self._doc_idx: List[int] = [0]

    @property
    def dtype(self):
        return self._dtype

    def add_item(self, arr):
        arr = np.asarray(arr, dtype=self._dtype)
        self._data_file.write(arr.tobytes(order="C"))
        self._sizes.append(arr.size)

    def end_document(self):
        self._doc_idx.append(len(self._sizes))

    def merge_file_(self, another_file: str):
        index = _MMapIndex(index_file_path(another_file), skip_warmup=True)
        assert index.dtype == self._dtype
        self._sizes.extend(int(s) for s in index.sizes)
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self._data_file)

    def finalize(self, index_file: str):
        self._data_file.close()
        _MMapIndex.write(index_file, self._dtype, self._sizes, self._doc_idx)
