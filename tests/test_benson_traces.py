"""Trace replay of REAL Benson runs (BASELINE configs 1-2): the poly__* call sequences the
unmodified bensolve CLI issued on the reference's own ex/*.vlp problems (recorded by
tests/golden/make_benson_traces.py; ex11 is the degenerate-branch stress) are
fed to every engine; return codes and final counts must equal the recorded reference run and the
canonical state must equal the checker's bit for bit."""
import pytest

from helpers import benson_files, check_benson_fixture

IDS = lambda p: p.split("/")[-1][:-5]
# syn_q5 (BASELINE config 4 shape, scaled): 4804 cuts in R^5 with recession directions on thousands of facets -- 25-45 s
# per engine on the CPU, so it runs once per suite (host double vs restatement here, product vs restatement on the GPU)
HEAVY = "syn_q5"
light_files = lambda: [p for p in benson_files() if HEAVY not in p]
FLAG_EAGER_GC, FLAG_MULTI_KERNEL, FLAG_TAIL_PHASES, FLAG_FORCE_WIDE = 2, 4, 8, 16


@pytest.mark.parametrize("path", light_files(), ids=IDS)
def test_oracle_on_benson_traces(oracle_lib, ref_lib, path):
    check_benson_fixture(oracle_lib, path, checker=ref_lib)


@pytest.mark.parametrize("path", benson_files(), ids=IDS)
def test_host_logic_on_benson_traces(emul_lib, oracle_lib, path):
    check_benson_fixture(emul_lib, path, checker=oracle_lib)


@pytest.mark.parametrize("path", [p for p in light_files() if "ex10" not in p], ids=IDS)
def test_host_logic_tail_phases_on_benson_traces(emul_lib, oracle_lib, path):
    check_benson_fixture(emul_lib, path, checker=oracle_lib, flags=FLAG_TAIL_PHASES | FLAG_EAGER_GC)


@pytest.mark.gpu
@pytest.mark.parametrize("path", benson_files(), ids=IDS)
def test_gpu_on_benson_traces(product_lib, oracle_lib, path):
    check_benson_fixture(product_lib, path, checker=oracle_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in benson_files() if "ex11" in p or "ex07" in p], ids=IDS)
def test_gpu_multi_kernel_path_on_benson_traces(product_lib, oracle_lib, path):
    check_benson_fixture(product_lib, path, checker=oracle_lib, flags=FLAG_MULTI_KERNEL | FLAG_EAGER_GC)


@pytest.mark.gpu
@pytest.mark.parametrize("path", light_files(), ids=IDS)
def test_gpu_forced_wide_cluster_on_benson_traces(product_lib, oracle_lib, path):
    """Every cut of the real Benson runs through the kernel variants the bench uses (k_tail<16>, grid-wide K4, k_tail2<16>)."""
    check_benson_fixture(product_lib, path, checker=oracle_lib, flags=FLAG_FORCE_WIDE)
