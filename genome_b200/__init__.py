"""genome_b200 -- B200-native k-mer -> de Bruijn graph path behind winger/genome's DNAMap / Graph surface.

The product is the C-ABI library libgenome_b200.so (include/genome_b200.h, sources in genome_b200/csrc).  The
Python modules here are the host-side harness used by tests and bench.py in place of the Scala shim
(INTEGRATION.md): `dnamap` mirrors trait DNAMap / PartitionedDNAMap / FreqFilter, `graph` mirrors trait Graph /
object Graph, `synth` generates read sets in the reference's `.bin` layout.
"""
from .dnamap import ArrayDNAMap, PartitionedDNAMap, FreqFilter, PairedEndData  # noqa: F401
from .graph import Graph, MapGraph  # noqa: F401
