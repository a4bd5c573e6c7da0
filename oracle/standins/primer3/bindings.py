"""``primer3.bindings.DEFAULT_P3_ARGS`` stand-in (``bin/Parameters.py:5,597-602``;
values confirmed by the reference's ``README.md:115-120``)."""
from types import SimpleNamespace

DEFAULT_P3_ARGS = SimpleNamespace(mv_conc=50.0, dv_conc=1.5, dntp_conc=0.6, dna_conc=50.0,
                                  temp_c=37.0, max_loop=30)
