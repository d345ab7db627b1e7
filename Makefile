# Developer convenience.  `lib` runs the contract entry point (__graft_entry__.build(): plan.cu + the per-length-group
# fast_inst.cu units, compiled in parallel and linked into hipgp_b200/csrc/libhipgp_b200.so); `dev` is the same build with
# the reduced length list (HIPGP_DEV_SMALL); `emu` builds the test-only CPU emulation.
SRC = hipgp_b200/csrc
lib:
	python -c "import __graft_entry__ as g; g._compile_lib()"
dev:
	HIPGP_DEV_SMALL=1 python -c "import __graft_entry__ as g; g._compile_lib()"
emu:
	python -c "import sys; sys.path.insert(0,'tests'); import emu_build; print(emu_build.build())"
.PHONY: lib dev emu
