"""Dataset loaders of the reference, kept and re-hosted without TensorFlow (``qa_cpg/data.py``).

Same class names, constructor arguments and public attributes (``dataset_name``, ``num_ent``, ``num_rel``,
``train_dataset(...)``, ``eval_dataset(...)``) as ``qa_cpg/data.py:25-698``; what changes is the storage and the
pipeline underneath:

* the reference turns ``train / valid|dev / test .txt`` into four JSON files, assigns ids, writes TFRecords and reads them
  back through ``tf.data`` (``data.py:341-475,574-594``).  Here the text files are parsed once into integer CSR arrays
  (one ``(e1, rel) -> {e2}`` row per query) cached as ``<directory>/coper_cache_<dataset>.npz``;
* a training sample is one ``(e1, rel)`` query with its whole positive set, reverse relations included
  (``_write_graph`` with ``labels=None``, ``data.py:482-488``); an evaluation sample is one ``(e1, rel, e2)`` triple
  whose filter set is every known true tail in train+dev+test (``data.py:464-467,494``); dev / test never get reverse
  edges (``add_reverse_per_filetype=[True, False, False]``, ``data.py:601,612``);
* batches are dicts of NumPy arrays in the schema ``models.ConvE.stage_batch`` accepts: int64 ``e1``, ``e2``, ``rel``
  ``[B]`` and ``e2_multi`` as CSR id lists (``e2_multi_rowptr`` int32 ``[B+1]``, ``e2_multi_col`` int32 ``[nnz]``) — the
  form the TFRecords hold (``data.py:574-594``) — instead of the dense fp32 multi-hot ``[B, N]`` the reference builds
  on CPU threads (``data.py:182-186,318-322``).  ``dense=True`` reproduces the reference's dense schema;
* ids: an existing ``entities.txt`` / ``relations.txt`` is honoured (``data.py:521-533``); otherwise ids are assigned in
  SORTED name order — the reference iterates Python sets, so its ids depend on ``PYTHONHASHSEED``;
* nothing is downloaded: there is no network here, so missing files raise ``FileNotFoundError`` naming the URL the
  reference would fetch (``data.py:58-72``).

Training labels: full 1-N (``num_labels=None``, ``_add_lookup_values``, ``data.py:314-330``) or the two sampled-label
pipelines the shipped configs use — ``_sample_negatives`` (``data.py:228-277``) and, with
``one_positive_label_per_sample``, ``_create_negative_sampling_dataset`` (``data.py:279-312``) — restated in NumPy: the
same index/label construction and the same distributions (TF's RNG streams cannot be reproduced; the "prefix / window
of a random permutation of all entities" the reference takes as negatives is drawn directly as a uniformly random
ordered set of distinct entities).  Sampled batches carry ``lookup_values`` int32 ``[B, L]`` and ``e2_multi`` fp32
``[B, L]`` exactly as ``models.py:135-152,165`` expects them.
"""
from __future__ import annotations

import logging
import os
import tarfile
from typing import Dict, Iterator, List, Optional

import numpy as np

__all__ = ["Loader", "NationsLoader", "UMLSLoader", "KinshipLoader", "WN18RRLoader", "YAGO310Loader",
           "FB15k237Loader", "CountriesS1Loader", "CountriesS2Loader", "CountriesS3Loader", "WN18Loader",
           "FB15kLoader", "NELL995Loader"]

logger = logging.getLogger(__name__)


class Loader:
    def __init__(self, url, filenames, dataset_name):
        self.url = url
        self.filenames = filenames
        self.dataset_name = dataset_name

    def load_and_preprocess(self, directory, buffer_size=1024 * 1024):
        raise NotImplementedError

    def maybe_download(self, directory, buffer_size=1024 * 1024):
        """The reference downloads missing files (``data.py:58-72``); offline we can only check for them."""
        missing = [f for f in self.filenames if not os.path.exists(os.path.join(directory, f))]
        if missing:
            raise FileNotFoundError(
                "dataset '%s': %s not found under %s and there is no network access; the reference would fetch "
                "them from %s" % (self.dataset_name, ", ".join(missing), directory, self.url))

    def maybe_extract(self, directory, buffer_size=1024 * 1024):
        self.maybe_download(directory, buffer_size)
        extracted = False
        for filename in self.filenames:
            if filename.endswith(".tar.gz"):
                path = os.path.join(directory, filename)
                target = path[:-7]
                if not os.path.exists(target):
                    logger.info("Extracting file: %s", path)
                    with tarfile.open(path, "r:*") as handle:
                        handle.extractall(path=target)
                extracted = True
        return extracted


class _Split:
    """CSR view of one split: query q = (e1[q], rel[q]) with tails col[rowptr[q]:rowptr[q+1]]."""

    def __init__(self, e1, rel, rowptr, col, is_inverse):
        self.e1, self.rel, self.rowptr, self.col, self.is_inverse = e1, rel, rowptr, col, is_inverse

    def __len__(self):
        return len(self.e1)


class _DataLoader(Loader):
    def __init__(self, url, filenames, dataset_name, filetypes=("train", "dev", "test"),
                 needs_test_set_cleaning=False, add_reverse_per_filetype=None):
        self.filetypes = list(filetypes)
        self.add_reverse_per_filetype = list(add_reverse_per_filetype or [True] * len(self.filetypes))
        self.needs_test_set_cleaning = needs_test_set_cleaning
        super().__init__(url, filenames, dataset_name)
        self.num_ent = None
        self.num_rel = None
        self._cache = None

    # ------------------------------------------------------------------------------------------ preprocessing
    def _read_triples(self, directory) -> Dict[str, List[tuple]]:
        if all(os.path.exists(os.path.join(directory, "%s.txt" % ft)) for ft in self.filetypes):
            pass                                  # the split files are already here: no archive needed
        elif self.maybe_extract(directory):
            directory = os.path.join(directory, self.dataset_name)      # the archive adds one directory level
        out = {}
        for ft in self.filetypes:
            rows = []
            with open(os.path.join(directory, "%s.txt" % ft), "r") as handle:
                for line in handle:
                    if not line.strip():
                        continue
                    e1, rel, e2 = (t.strip() for t in line.split("\t"))
                    rows.append((e1, rel, e2))
            out[ft] = rows
        return out, directory

    def load_and_preprocess(self, directory, buffer_size=1024 * 1024):
        """Parse the splits into id-space graphs (``data.py:401-475``) and cache them.  Returns the cache dict."""
        if self._cache is not None:
            return self._cache
        cache_file = os.path.join(directory, "coper_cache_%s%s.npz" % (
            self.dataset_name, "_clean" if self.needs_test_set_cleaning else ""))
        if os.path.exists(cache_file):
            try:
                z = np.load(cache_file, allow_pickle=False)
                cache = {k: z[k] for k in z.files}
                self.num_ent, self.num_rel = int(cache["num_ent"]), int(cache["num_rel"])
                self._cache = cache
                return self._cache
            except Exception as exc:              # truncated / corrupt file (e.g. an interrupted writer): rebuild it
                logger.warning("ignoring unreadable cache %s (%r); preprocessing again", cache_file, exc)
        logger.info("Loading and preprocessing the '%s' dataset.", self.dataset_name)
        triples, data_dir = self._read_triples(directory)
        # graphs keyed by (e1, rel) -> set(e2); reverse edges per split as in data.py:427-439
        full, graphs = {}, {ft: {} for ft in self.filetypes}
        for i, ft in enumerate(self.filetypes):
            g = graphs[ft]
            for e1, rel, e2 in triples[ft]:
                rev = rel + "_reverse"
                full.setdefault((e1, rel), set()).add(e2)
                full.setdefault((e2, rev), set()).add(e1)
                g.setdefault((e1, rel), set()).add(e2)
                g.setdefault((e2, rev), set())
                if self.add_reverse_per_filetype[i]:
                    g[(e2, rev)].add(e1)
        allowed_e = allowed_r = None
        if self.needs_test_set_cleaning:                              # data.py:448-461
            allowed_e, allowed_r = set(), set()
            for (e1, rel), tails in graphs[self.filetypes[0]].items():
                allowed_e.add(e1)
                allowed_e.update(tails)
                allowed_r.add(rel)
        ent_ids, rel_ids = self._assign_ids(data_dir, full, allowed_e, allowed_r)
        self.num_ent, self.num_rel = len(ent_ids), len(rel_ids)
        cache = {"num_ent": np.int64(self.num_ent), "num_rel": np.int64(self.num_rel)}

        def csr(keys_tails):
            e1 = np.array([ent_ids[k[0]] for k, _ in keys_tails], np.int64)
            rel = np.array([rel_ids[k[1]] for k, _ in keys_tails], np.int64)
            inv = np.array([k[1].endswith("_reverse") for k, _ in keys_tails], bool)
            rowptr = np.zeros(len(keys_tails) + 1, np.int64)
            rowptr[1:] = np.cumsum([len(t) for _, t in keys_tails])
            col = np.fromiter((ent_ids[e] for _, t in keys_tails for e in sorted(t)), np.int32, int(rowptr[-1]))
            return e1, rel, rowptr, col, inv

        # train: one sample per (e1, rel) key with its own positives (data.py:482-488); keys with no tails (reverse
        # keys of a split without reverse edges) carry an empty label set exactly like the reference's records
        train = sorted(graphs[self.filetypes[0]].items())
        for nm, arr in zip(("e1", "rel", "rowptr", "col", "inv"), csr(train)):
            cache["train_" + nm] = arr
        # eval splits: one sample per (e1, rel, e2) with the FULL-graph label set as filter (data.py:489-503)
        for ft, out_name in zip(self.filetypes, ("train", "dev", "test")):
            e1s, rels, e2s, keys = [], [], [], []
            for (e1, rel), tails in sorted(graphs[ft].items()):
                if allowed_e is not None and e1 not in allowed_e:
                    continue
                if allowed_r is not None and rel not in allowed_r:
                    continue
                for e2 in sorted(tails):
                    if allowed_e is not None and e2 not in allowed_e:
                        continue
                    e1s.append(ent_ids[e1]); rels.append(rel_ids[rel]); e2s.append(ent_ids[e2])
                    keys.append((e1, rel))
            uniq = sorted(set(keys))
            index = {k: i for i, k in enumerate(uniq)}
            fe1, frel, frowptr, fcol, finv = csr([(k, full[k]) for k in uniq])
            cache["eval_%s_e1" % out_name] = np.array(e1s, np.int64)
            cache["eval_%s_rel" % out_name] = np.array(rels, np.int64)
            cache["eval_%s_e2" % out_name] = np.array(e2s, np.int64)
            cache["eval_%s_key" % out_name] = np.array([index[k] for k in keys], np.int64)
            cache["eval_%s_inv" % out_name] = np.array([k[1].endswith("_reverse") for k in keys], bool)
            cache["eval_%s_rowptr" % out_name] = frowptr
            cache["eval_%s_col" % out_name] = fcol
        try:                                                          # temp file + rename: readers never see half a file
            tmp = "%s.tmp%d.npz" % (cache_file, os.getpid())
            np.savez(tmp, **cache)
            os.replace(tmp, cache_file)
        except OSError:                                               # read-only dataset directory: keep in memory
            logger.warning("could not write %s; keeping the preprocessed graphs in memory", cache_file)
        self._cache = cache
        return cache

    @staticmethod
    def _assign_ids(directory, full, allowed_e=None, allowed_r=None):
        """``entities.txt`` / ``relations.txt`` if present (``data.py:521-533``), else sorted names (written back when
        the directory is writable).  Relation ids include the ``_reverse`` relations."""
        def load(path):
            with open(path, "r") as handle:
                return {line.strip(): i for i, line in enumerate(handle) if line.strip()}
        ent_file, rel_file = os.path.join(directory, "entities.txt"), os.path.join(directory, "relations.txt")
        if os.path.exists(ent_file):
            ent_ids = load(ent_file)
        else:
            # the reference numbers what its 'full' JSON mentions: every kept query (e1 and rel allowed) with its
            # whole full-graph label set (data.py:489-503,536-552)
            names = set()
            for (e1, rel), tails in full.items():
                if allowed_r is not None and rel not in allowed_r:
                    continue
                if allowed_e is not None and e1 not in allowed_e:
                    continue
                names.add(e1)
                names.update(tails)
            ent_ids = {n: i for i, n in enumerate(sorted(names))}
        if os.path.exists(rel_file):
            rel_ids = load(rel_file)
        else:
            rels = sorted({rel for (_, rel) in full if allowed_r is None or rel in allowed_r})
            rel_ids = {n: i for i, n in enumerate(rels)}
        for path, ids in ((ent_file, ent_ids), (rel_file, rel_ids)):
            if not os.path.exists(path):
                try:
                    tmp = "%s.tmp%d" % (path, os.getpid())
                    with open(tmp, "w") as handle:
                        for name, _ in sorted(ids.items(), key=lambda kv: kv[1]):
                            handle.write(name + "\n")
                    os.replace(tmp, path)
                except OSError:
                    pass
        return ent_ids, rel_ids

    # kept for API parity with data.py:332-399 (ids become known here, as in the reference)
    def generate_json_files_and_ids(self, directory, buffer_size=1024 * 1024):
        return self.load_and_preprocess(directory, buffer_size)

    def maybe_create_tf_record_files(self, directory, max_records_per_file=1000000, buffer_size=1024 * 1024):
        """The reference materialises TFRecords here; we materialise the CSR cache.  Sets num_ent / num_rel."""
        return self.load_and_preprocess(directory, buffer_size)

    # ------------------------------------------------------------------------------------------ pipelines
    def train_dataset(self, directory, batch_size, include_inv_relations=True, num_parallel_readers=32,
                      num_parallel_batches=32, buffer_size=1024 * 1024, prefetch_buffer_size=10, prop_negatives=10.0,
                      num_labels=None, cache=False, one_positive_label_per_sample=True, seed=0,
                      dense=False, device_sampling=False) -> Iterator[Dict[str, np.ndarray]]:
        """Endless iterator of training batches (``data.py:89-166``: repeat -> labels -> shuffle -> batch).
        The reference shuffles with a 1000-element buffer; here every epoch is a seeded full permutation.
        ``device_sampling`` (with ``num_labels``, not with ``one_positive_label_per_sample``): batches carry the CSR id
        lists plus ``sample_on_device = (num_labels, prop_negatives)``; ``ConvE.train_step`` then draws the
        ``[B, num_labels]`` ids / labels of ``_sample_negatives`` (``data.py:228-277``) on the GPU."""
        c = self.load_and_preprocess(directory, buffer_size)
        on_device = None
        if num_labels is not None and device_sampling and not one_positive_label_per_sample:
            if num_labels > self.num_ent:
                raise ValueError("Parameter `num_labels` needs to be at most the total number of entities.")
            on_device = (int(num_labels), float(prop_negatives))
            num_labels = None
        if num_labels is not None:
            if num_labels > self.num_ent:                         # data.py:149
                raise ValueError("Parameter `num_labels` needs to be at most the total number of entities.")
            return self._sampled_train_dataset(c, batch_size, include_inv_relations, prop_negatives, int(num_labels),
                                               one_positive_label_per_sample, seed, prefetch_buffer_size)
        sel = np.arange(len(c["train_e1"]))
        if not include_inv_relations:
            sel = sel[~c["train_inv"]]
        split = _Split(c["train_e1"], c["train_rel"], c["train_rowptr"], c["train_col"], c["train_inv"])
        rng = np.random.default_rng(seed)

        def gen():
            while True:
                order = rng.permutation(sel)
                for s in range(0, len(order), batch_size):
                    idx = order[s:s + batch_size]
                    bt = self._make_batch(split.e1[idx], split.rel[idx], np.full(len(idx), -1, np.int64),
                                          idx, split.rowptr, split.col, dense)
                    if on_device is not None:
                        bt["sample_on_device"] = on_device
                    yield bt
        return gen()

    # ------------------------------------------------------------------------------------------ sampled labels
    @staticmethod
    def _distinct_random(rng, num_ent, rows, k):
        """[rows, k] uniformly random DISTINCT entity ids per row, in random order (== the first k entries of an
        independent random permutation per row, data.py:238,270-272), by vectorised rejection of repeats."""
        x = rng.integers(0, num_ent, (rows, k), dtype=np.int64)
        if k <= 1:
            return x
        while True:
            order = np.argsort(x, axis=1, kind="stable")
            xs = np.take_along_axis(x, order, axis=1)
            dup_sorted = np.zeros(x.shape, bool)
            dup_sorted[:, 1:] = xs[:, 1:] == xs[:, :-1]
            n = int(dup_sorted.sum())
            if n == 0:
                return x
            dup = np.zeros(x.shape, bool)
            np.put_along_axis(dup, order, dup_sorted, axis=1)
            x[dup] = rng.integers(0, num_ent, n, dtype=np.int64)

    def _sampled_train_dataset(self, c, batch_size, include_inv, prop_negatives, num_labels, one_pos, seed, prefetch):
        e1a, rela, rowptr, col, inv = (c["train_" + k] for k in ("e1", "rel", "rowptr", "col", "inv"))
        sel = np.arange(len(e1a))
        if not include_inv:
            sel = sel[~inv]
        sel = sel[(rowptr[sel + 1] - rowptr[sel]) > 0]              # a query without positives yields no sampled row
        N = self.num_ent
        rng = np.random.default_rng(seed)
        n_pos_needed = int(1.0 / (1.0 + prop_negatives) * num_labels)     # data.py:244

        def labels_of(rows, lookup):
            """label 1 where the looked-up entity is a positive of its query (tf.gather(e2s_dense, indexes))."""
            out = np.zeros(lookup.shape, np.float32)
            for i, r in enumerate(rows):
                out[i] = np.isin(lookup[i], col[rowptr[r]:rowptr[r + 1]])
            return out

        def batch_sample_negatives(rows):                            # data.py:228-277
            B = len(rows)
            lookup = np.empty((B, num_labels), np.int64)
            wrong = self._distinct_random(rng, N, B, num_labels)     # prefix of a random permutation of all entities
            for i, r in enumerate(rows):
                pos = rng.permutation(col[rowptr[r]:rowptr[r + 1]])
                if len(pos) <= n_pos_needed:
                    n_pos = len(pos)
                else:
                    n_pos = num_labels - min(N, num_labels - n_pos_needed)
                lookup[i, :n_pos] = pos[:n_pos]
                lookup[i, n_pos:] = wrong[i, :num_labels - n_pos]
            return rows, lookup

        def gen_rows_one_positive(order):                            # data.py:279-312: one row per positive
            for r in order:
                pos = col[rowptr[r]:rowptr[r + 1]]
                for p in pos:
                    yield r, p

        def rows_one_positive_forever():
            """repeat() comes before the per-positive expansion in the reference (data.py:147): an endless row stream"""
            while True:
                yield from gen_rows_one_positive(rng.permutation(sel))

        def shuffled(stream, buffer_size=1000):
            """tf.data shuffle(buffer_size) (data.py:160): a row leaves the buffer at a uniformly random position, so
            the consecutive positives of one high-degree query are spread over ~buffer_size rows instead of filling a
            batch"""
            buf = []
            for item in stream:
                if len(buf) < buffer_size:
                    buf.append(item)
                    continue
                j = int(rng.integers(0, buffer_size))
                out, buf[j] = buf[j], item
                yield out

        def gen():
            if one_pos:
                buf_r, buf_p = [], []
                for r, p in shuffled(rows_one_positive_forever()):     # never truncated at an epoch boundary
                    buf_r.append(r)
                    buf_p.append(p)
                    if len(buf_r) == batch_size:
                        rows = np.array(buf_r)
                        lookup = np.concatenate([np.array(buf_p, np.int64)[:, None],
                                                 self._distinct_random(rng, N, len(rows), num_labels - 1)], axis=1)
                        yield self._sampled_batch(e1a, rela, rows, lookup, labels_of(rows, lookup))
                        buf_r, buf_p = [], []
            while True:
                order = rng.permutation(sel)
                for s0 in range(0, len(order), batch_size):
                    rows, lookup = batch_sample_negatives(order[s0:s0 + batch_size])
                    yield self._sampled_batch(e1a, rela, rows, lookup, labels_of(rows, lookup))
        return _prefetch(gen(), prefetch)

    @staticmethod
    def _sampled_batch(e1a, rela, rows, lookup, labels):
        return {"e1": e1a[rows].astype(np.int64), "rel": rela[rows].astype(np.int64),
                "e2": np.full(len(rows), -1, np.int64), "e2_multi": labels.astype(np.float32),
                "lookup_values": lookup.astype(np.int32)}

    def eval_dataset(self, directory, dataset_type, batch_size, include_inv_relations=True, buffer_size=1024 * 1024,
                     prefetch_buffer_size=10, dense=False) -> Iterator[Dict[str, np.ndarray]]:
        """One pass over the split (``data.py:168-226``); exhaustion plays the role of tf.errors.OutOfRangeError."""
        c = self.load_and_preprocess(directory, buffer_size)
        p = "eval_%s_" % dataset_type
        sel = np.arange(len(c[p + "e1"]))
        if not include_inv_relations:
            sel = sel[~c[p + "inv"]]

        def gen():
            for s in range(0, len(sel), batch_size):
                idx = sel[s:s + batch_size]
                yield self._make_batch(c[p + "e1"][idx], c[p + "rel"][idx], c[p + "e2"][idx], c[p + "key"][idx],
                                       c[p + "rowptr"], c[p + "col"], dense)
        return gen()

    def _make_batch(self, e1, rel, e2, rows, rowptr, col, dense):
        lens = rowptr[rows + 1] - rowptr[rows]
        out_ptr = np.zeros(len(rows) + 1, np.int32)
        out_ptr[1:] = np.cumsum(lens)
        out_col = np.empty(int(out_ptr[-1]), np.int32)
        for i, r in enumerate(rows):
            out_col[out_ptr[i]:out_ptr[i + 1]] = col[rowptr[r]:rowptr[r + 1]]
        batch = {"e1": e1.astype(np.int64), "rel": rel.astype(np.int64), "e2": e2.astype(np.int64),
                 "lookup_values": np.zeros((len(rows), 0), np.int32)}          # data.py:206-213,314-330
        if dense:
            m = np.zeros((len(rows), self.num_ent), np.float32)
            m[np.repeat(np.arange(len(rows)), lens), out_col] = 1.0
            batch["e2_multi"] = m
        else:
            batch["e2_multi_rowptr"], batch["e2_multi_col"] = out_ptr, out_col
        return batch


def _prefetch(iterator, depth):
    """tf.data's .prefetch(n): a daemon thread keeps up to n batches ready while the device works."""
    if not depth or depth <= 0:
        return iterator
    import queue
    import threading
    q = queue.Queue(maxsize=int(depth))

    def worker():
        try:
            for item in iterator:
                q.put(item)
        except BaseException as exc:          # surface producer errors in the consumer
            q.put(exc)
    threading.Thread(target=worker, daemon=True).start()

    def consume():
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            yield item
    return consume()


class _ConvEDataLoader(_DataLoader):
    def __init__(self, dataset_name, needs_test_set_cleaning=False):
        super().__init__("https://github.com/TimDettmers/ConvE/raw/master", [dataset_name + ".tar.gz"], dataset_name,
                         ["train", "valid", "test"], needs_test_set_cleaning, [True, False, False])


class _MinervaDataLoader(_DataLoader):
    def __init__(self, dataset_name, needs_test_set_cleaning=False):
        super().__init__("https://raw.githubusercontent.com/shehzaadzd/MINERVA/master/datasets/data_preprocessed/%s"
                         % dataset_name, ["train.txt", "dev.txt", "test.txt"], dataset_name,
                         ["train", "dev", "test"], needs_test_set_cleaning, [True, False, False])


class NationsLoader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("nations")


class UMLSLoader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("umls")


class KinshipLoader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("kinship")


class WN18RRLoader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("WN18RR")


class YAGO310Loader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("YAGO3-10")


class FB15k237Loader(_ConvEDataLoader):
    def __init__(self):
        super().__init__("FB15k-237")


class CountriesS1Loader(_MinervaDataLoader):
    def __init__(self):
        super().__init__("countries_S1")


class CountriesS2Loader(_MinervaDataLoader):
    def __init__(self):
        super().__init__("countries_S2")


class CountriesS3Loader(_MinervaDataLoader):
    def __init__(self):
        super().__init__("countries_S3")


class WN18Loader(_ConvEDataLoader):
    def __init__(self, is_test=False, needs_test_set_cleaning=False):
        self.is_test = is_test
        super().__init__("WN18" + ("-test" if is_test else ""), needs_test_set_cleaning)


class FB15kLoader(_ConvEDataLoader):
    def __init__(self, is_test=False, needs_test_set_cleaning=False):
        self.is_test = is_test
        super().__init__("FB15k" + ("test" if is_test else ""), needs_test_set_cleaning)


class NELL995Loader(_MinervaDataLoader):
    def __init__(self, is_test=False, needs_test_set_cleaning=False):
        self.is_test = is_test
        # NELL contains test entities that never appear in training; the reference removes them (data.py:690-698)
        super().__init__("nell-995" + ("-test" if is_test else ""), needs_test_set_cleaning)
