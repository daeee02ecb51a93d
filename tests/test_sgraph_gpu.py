"""The sharded Graph.buildGraph (csrc/sgraph.cuh) on the device: P virtual ranks on one GPU against the oracle and against
the single-GPU build.  The same functors and orchestration pass tests/test_sgraph_emul_cpu.py through a g++ backend; these
tests cover the CUDA backend (launches, atomics, scans, arena memory)."""
import os

import numpy as np
import pytest

from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap, FreqFilter, PairedEndData
from genome_b200.graph import Graph
from oracle import pyoracle
from tests import helpers as H
from tests.test_sgraph_emul_cpu import GRAPH_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", GRAPH_CASES)
def test_virtual_shards_match_oracle(gpu, k, glen, rl, cov, err, rounds):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, rounds)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(rounds)
    og = pyoracle.OracleGraph(om)
    for P in (1, 2, 3, 8, 16):
        g = Graph.buildGraphVirtualShards(k, gm, P)
        g.check()
        assert g.counts() == og.counts(), P
        H.assert_graph_equal(g, og)
        g.close()
    # the operators run on the result like on any graph
    g = Graph.buildGraphVirtualShards(k, gm, 4)
    g.retain_largest(); og.retain_largest()
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)


def test_virtual_shards_noncanonical_and_cycle(gpu):
    k = 9
    rng = np.random.default_rng(3)
    genome = synth.random_genome(2000, 99)
    fw = np.array([synth.kmer_to_int(synth.decode(genome[i:i + k])) for i in range(genome.size - k + 1)], np.uint64)
    rc = np.array([pyoracle.revcomp(int(x), k) for x in fw], np.uint64)
    pick = rng.random(fw.size)
    keys = np.concatenate([fw[pick < 0.6], rc[pick > 0.4]])
    gm = ArrayDNAMap(k)
    om = pyoracle.OracleMap(k)
    gm.update_counts(keys)
    for x in keys:
        om.update1(int(x))
    og = pyoracle.OracleGraph(om)
    for P in (1, 4, 8):
        g = Graph.buildGraphVirtualShards(k, gm, P)
        H.assert_graph_equal(g, og)
    # a perfect cycle that runs through every rank: nothing is built, every oriented vertex is reported as a cycle vertex
    k = 11
    genome = synth.random_genome(500, 8)
    circ = np.concatenate([genome, genome[:k - 1]])
    keys = np.array([pyoracle.canonical(synth.kmer_to_int(synth.decode(circ[i:i + k])), k) for i in range(genome.size)], np.uint64)
    gm = ArrayDNAMap(k)
    gm.update_counts(keys)
    for P in (1, 8):
        g = Graph.buildGraphVirtualShards(k, gm, P)
        assert g.counts() == (0, 0, 0)
        assert g.stats()["cycle_vertices"] == 2 * genome.size


def test_virtual_shards_equal_single_gpu_build_at_size(gpu):
    """300 kbp error-free genome: 4 nodes / 2 edges whose ~300 k interior vertices cross the 8 ranks ~25 k times; and a noisy
    1 Mbp-read set compared with gb_graph_build edge for edge."""
    k = 31
    b, n, genome = H.small_reads(300000, 100, 40, 0.0, seed=42)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 1)
    g1 = Graph.buildGraph(k, gm)
    g8 = Graph.buildGraphVirtualShards(k, gm, 8)
    assert g8.counts() == g1.counts() and g8.counts()[:2] == (4, 2)
    assert H.canon_gpu_graph(g8) == H.canon_gpu_graph(g1)
    b, n, _ = H.small_reads(400000, 100, 25, 0.01, seed=43)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 2)
    g1 = Graph.buildGraph(k, gm)
    for P in (2, 8):
        gp = Graph.buildGraphVirtualShards(k, gm, P)
        assert gp.counts() == g1.counts()
        assert H.canon_gpu_graph(gp) == H.canon_gpu_graph(g1)
        assert gp.stats()["cycle_vertices"] == g1.stats()["cycle_vertices"]


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7])
def test_random_dense_kmer_sets(gpu, k):
    """Random subsets of the whole k-mer space for small k (tests/test_sgraph_emul_cpu.py has the same sets): self-loops,
    hairpins, palindromes, perfect cycles; through BOTH device builds (gb_graph_build and the virtual shards)."""
    rng = np.random.default_rng(700 + k)
    space = 1 << (2 * k)
    for trial in range(12):
        frac = [0.05, 0.2, 0.5, 0.9][trial % 4]
        xs = np.flatnonzero(rng.random(space) < frac).astype(np.uint64)
        if xs.size == 0:
            continue
        dual = trial % 3 == 2
        keys = xs if dual else np.array([pyoracle.canonical(int(x), k) for x in xs], np.uint64)
        om = pyoracle.OracleMap(k)
        for x in keys.tolist():
            om.update1(x)
        og = pyoracle.OracleGraph(om)
        gm = ArrayDNAMap(k)
        gm.update_counts(keys)
        assert gm.size == om.size()
        P = int(rng.integers(1, 9))
        for g in (Graph.buildGraph(k, gm), Graph.buildGraphVirtualShards(k, gm, P)):
            assert g.counts() == og.counts(), (k, trial, P, dual)
            H.assert_graph_equal(g, og)
            g.close()
        gm.close()
