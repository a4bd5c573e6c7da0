"""Minimal stand-in for ``biopython`` (absent here).  TEST INFRASTRUCTURE ONLY.
Covers ``Bio.Seq.Seq``, ``Bio.SeqRecord.SeqRecord`` and ``Bio.SeqIO.parse`` for
'fasta' and 'genbank' as used by ``bin/getCandidateKmers.py:1,8,93-95,212-217``
and ``bin/Primer.py:3,87``."""
__version__ = "1.81"
