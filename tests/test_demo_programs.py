"""The reference's own demo programs as black-box callers.  oracle/_ref/tri_demo and bfs_demo are
compiled, unmodified, from Demo/Program/*.c (oracle/Makefile.ref); the expected numbers are the ones
the reference prints in Demo/Output/tri_demo.out and bfs_demo.out for the inputs that the programs
generate themselves (Demo/demo:39-79), so no input file is needed.

CPU: the programs run against the plain reference library (pins the build recipe and the known
answers).  GPU: the SAME binaries run with LD_PRELOAD=libgb_b200_shim.so -- nothing is relinked -- and
must print the same numbers, with the shim reporting that the multiplies ran on the device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
SHIM = os.path.join(ROOT, "graphblas_b200", "libgb_b200_shim.so")

# args -> "# triangles" (Demo/Output/tri_demo.out:66,125,1132,1218,1304,1391)
TRI = {("1", "4", "4"): 872, ("0", "5", "5", "30", "1"): 7, ("1", "200", "200", "0"): 2160400,
       ("0", "10000", "10000", "100000", "0"): 1357, ("0", "10000", "10000", "100000", "1"): 1357}
TRI_BIG = {("0", "100000", "100000", "10000000", "0"): 1330131}
# args -> (nodes reachable from node 0, max BFS level) (Demo/Output/bfs_demo.out:61-62,116-117,1057-1058,1136-1137)
BFS = {("1", "4", "4"): (65, 5), ("0", "5", "5", "30", "1"): (5, 3), ("1", "200", "200", "0"): (120801, 201),
       ("0", "10000", "10000", "100000", "0"): (10000, 5)}


# args -> size of the maximal independent set (Demo/Output/mis_demo.out:59,94,765,827,952); the
# program verifies the set itself ("maximal independent set status verified")
MIS = {("1", "4", "4"): 13, ("0", "5", "5", "30", "1"): 1, ("1", "200", "200", "0"): 18430,
       ("0", "10000", "10000", "100000", "0"): 1684}
MIS_BIG = {("0", "100000", "100000", "10000000", "0"): 2798}


def run(prog, args, gpu):
    exe = os.path.join(REFDIR, prog)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (make -C oracle -f Makefile.ref)")
    env = dict(os.environ)
    env.pop("GB200_SHIM_DISABLE", None)
    if gpu:
        env["LD_PRELOAD"] = SHIM
        env["GB200_SHIM_STATS"] = "1"
    r = subprocess.run([exe, *args], capture_output=True, text=True, env=env, timeout=600, cwd=REFDIR)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


def tri_count(out):
    assert "error!" not in out                      # the program compares its two methods itself
    found = re.findall(r"^# triangles (\d+)", out, re.M)
    assert found, out[-500:]
    return int(found[0])


def bfs_result(out):
    reach = set(re.findall(r"nodes reachable from node 0: (\d+) out of", out))
    lev = set(re.findall(r"max BFS level: (\d+)", out))
    assert len(reach) == 1 and len(lev) == 1, (reach, lev)     # the four BFS variants agree
    return int(reach.pop()), int(lev.pop())


def mis_size(out):
    assert "error" not in out and "maximal independent set status verified" in out
    found = re.findall(r"independent set found: (\d+) of", out)
    assert found
    return int(found[0])


def gpu_calls(err):
    m = re.search(r"\[gb_b200 shim\] gpu_calls=(\d+) forwarded=(\d+) declined=(\d+)", err)
    assert m, "the shim was not loaded: " + err[-500:]
    return tuple(int(v) for v in m.groups())


@pytest.mark.parametrize("args", list(TRI))
def test_tri_demo_reference(args):
    out, _ = run("tri_demo", args, gpu=False)
    assert tri_count(out) == TRI[args]


@pytest.mark.parametrize("args", list(BFS))
def test_bfs_demo_reference(args):
    out, _ = run("bfs_demo", args, gpu=False)
    assert bfs_result(out) == BFS[args]


@pytest.mark.gpu
@pytest.mark.parametrize("args", list(TRI) + list(TRI_BIG))
def test_tri_demo_unmodified_binary_on_gpu(args):
    out, err = run("tri_demo", args, gpu=True)
    assert tri_count(out) == {**TRI, **TRI_BIG}[args]
    calls, forwarded, declined = gpu_calls(err)
    assert calls >= 2 and forwarded == 0 and declined == 0     # C<U>=L'*U (dot) and C<L>=L*L (saxpy)


@pytest.mark.gpu
@pytest.mark.parametrize("args", list(BFS))
def test_bfs_demo_unmodified_binary_on_gpu(args):
    out, err = run("bfs_demo", args, gpu=True)
    assert bfs_result(out) == BFS[args]
    calls, forwarded, declined = gpu_calls(err)
    assert calls >= 4 and forwarded == 0 and declined == 0     # one GrB_vxm per level and variant


@pytest.mark.parametrize("args", list(MIS))
def test_mis_demo_reference(args):
    """Luby's algorithm: GrB_vxm over MAX_FIRST_FP64 with a candidates mask, then LOR_LAND_BOOL"""
    out, _ = run("mis_demo", args, gpu=False)
    assert mis_size(out) == MIS[args]


@pytest.mark.gpu
@pytest.mark.parametrize("args", list(MIS) + list(MIS_BIG))
def test_mis_demo_unmodified_binary_on_gpu(args):
    out, err = run("mis_demo", args, gpu=True)
    assert mis_size(out) == {**MIS, **MIS_BIG}[args]
    calls, forwarded, declined = gpu_calls(err)
    assert calls >= 2 and forwarded == 0 and declined == 0
