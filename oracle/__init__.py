"""oracle/ -- TEST INFRASTRUCTURE ONLY.  CPU checkers for the x265 hot path:
   * liboracle.so       : our plain-C restatement (x265_oracle.c), cites reference file:line.
   * _ref/libx265ref*.so: the UNMODIFIED reference C sources compiled by oracle/Makefile
                          (kind "reference"); built in the authoring container, shipped prebuilt.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (libx265b200.so) never links or calls anything here."""
from .loader import orc, ref, have_ref, build  # noqa: F401
