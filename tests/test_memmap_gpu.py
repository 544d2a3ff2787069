"""the on-disk format.  `SearchArray.index(..., data_dir=...)` writes the posting words to a `.dat`
file and memory-maps them (reference phrase/memmap_arrays.py:145-208); the upload DMAs straight from that mapping
(cudaHostRegister) and a pickle carries the file name, not the words (reference test/test_search.py:62-73:
index with data_dir -> score -> pickle -> reload -> score)."""
import ctypes
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_memmap_pickle_round_trip(tmp_path):
    from searcharray_b200 import SearchArray
    docs = ["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny", ""] * 2000
    plain = SearchArray.index(docs)
    arr = SearchArray.index(docs, batch_size=1000, data_dir=str(tmp_path))
    assert isinstance(arr.host.words, np.memmap) and os.path.exists(arr.host.words_file)
    assert np.array_equal(arr.host.words, plain.host.words)
    want = plain.score("bar")
    want_ph = plain.score(["foo", "bar"])
    assert np.array_equal(arr.score("bar"), want) and np.array_equal(arr.score(["foo", "bar"]), want_ph)
    assert arr.score(["nope", "bar"]).sum() == 0
    with open(tmp_path / "data.pkl", "wb") as f:
        pickle.dump(arr, f)
    assert os.path.getsize(tmp_path / "data.pkl") < arr.host.words.nbytes        # the words did not travel
    with open(tmp_path / "data.pkl", "rb") as f:
        reloaded = pickle.load(f)
    assert isinstance(reloaded.host.words, np.memmap)
    assert np.array_equal(reloaded.score("bar"), want) and np.array_equal(reloaded.score(["foo", "bar"]), want_ph)
    assert np.array_equal(reloaded[1::2].termfreqs("bar"), plain.termfreqs("bar")[1::2])


def test_large_memmap_upload_is_page_locked_in_place(tmp_path):
    """a > 32 MB words file goes to HBM by DMA from the (read-only) file mapping, or through pinned bounce buffers"""
    from oracle import search as osearch
    from searcharray_b200 import SearchArray, _lib, synth
    spec = synth.SynthSpec(3_000_000, terms_per_bucket=3, n_phrases=8, n_bigrams=2)
    host, _, _ = synth.generate_shard(spec)
    host.avg_doc_length = synth.global_avg_doc_length(spec)
    assert host.words.nbytes > (32 << 20)
    host.memmap(str(tmp_path))
    arr = SearchArray.from_host_index(host)
    mode = ctypes.c_int(-1)
    _lib.check(_lib.lib().sa_index_upload_mode(arr._device().handle, ctypes.byref(mode)))
    assert mode.value in (1, 2), mode.value
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length, corpus_size=host.n_docs)
    for name in (spec.bucket_terms[0][0], spec.bucket_terms[3][1]):
        t = spec.term_index[name]
        assert np.array_equal(arr.score(name).view(np.uint32), oidx.score(t).view(np.uint32))
    ph = spec.phrases[0]["terms"]
    assert np.array_equal(arr.termfreqs(ph), oidx.termfreqs([spec.term_index[t] for t in ph]))
    # the same index from plain (anonymous) memory: registered in place as well
    host2, _, _ = synth.generate_shard(spec)
    arr2 = SearchArray.from_host_index(host2, avg_doc_length=host.avg_doc_length)
    _lib.check(_lib.lib().sa_index_upload_mode(arr2._device().handle, ctypes.byref(mode)))
    assert mode.value in (1, 2)
    assert np.array_equal(arr2.termfreqs(ph), arr.termfreqs(ph))
