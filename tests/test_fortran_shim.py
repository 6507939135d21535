"""The Fortran ISO_C_BINDING shim (fortran/*.f95) cannot be compiled here (no Fortran compiler in the image), so it is
held to the C ABI in three other ways: (1) every C entry point and struct field it binds exists in include/i3rc_b200.h
with the same spelling and order; (2) tests/ctest/shim_replay.c makes the shim's calls in the shim's order from plain C
and must reproduce the Python mirror's numbers on the GPU; (3) the Mersenne-Twister of the replacement module
RandomNumbers is checked by a statement-for-statement Python transliteration against the published known answers."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORTRAN = os.path.join(ROOT, "fortran")
HEADER = open(os.path.join(ROOT, "include", "i3rc_b200.h")).read()


def _src(name):
    return open(os.path.join(FORTRAN, name)).read()


def test_replacement_modules_keep_the_reference_public_lists():
    """Module names and public lists of Integrators/monteCarloRadiativeTransfer.f95:154-156,
    Code/monteCarloIllumination.f95:54-55 and Code/RandomNumbersForMC.f95:108-110."""
    want = {
        "monteCarloRadiativeTransfer_b200.f95": ("monteCarloRadiativeTransfer", [
            "integrator", "new_Integrator", "copy_Integrator", "isReady_Integrator", "finalize_Integrator", "specifyParameters",
            "computeRadiativeTransfer", "reportResults"]),
        "monteCarloIllumination_b200.f95": ("monteCarloIllumination", [
            "photonStream", "new_PhotonStream", "finalize_PhotonStream", "morePhotonsExist", "getNextPhoton"]),
        "RandomNumbersForMC_b200.f95": ("RandomNumbers", [
            "randomNumberSequence", "new_RandomNumberSequence", "finalize_RandomNumberSequence", "getRandomInt",
            "getRandomPositiveInt", "getRandomReal", "getRandomDouble"]),
    }
    for fname, (module, names) in want.items():
        s = _src(fname)
        assert re.search(rf"^module {module}\s*$", s, flags=re.M | re.I), fname
        public = " ".join(re.findall(r"public\s*::\s*((?:[^\n&]|&\s*\n)*)", s, flags=re.I)).replace("&", " ")
        listed = {t.strip().lower() for t in re.split(r"[,\s]+", public) if t.strip()}
        for n in names:
            assert n.lower() in listed, (fname, n)
    ill = _src("monteCarloIllumination_b200.f95")
    for ctor in ("newPhotonStream_Directional", "newPhotonStream_RandomAzimuth", "newPhotonStream_Flux", "newPhotonStream_Spotlight",
                 "newPhotonStream_Internal_Flux", "newPhotonStream_Internal_Intensity"):
        assert re.search(rf"function {ctor}\(", ill), ctor


def test_shim_binds_only_symbols_the_header_declares():
    s = _src("monteCarloRadiativeTransfer_b200.f95")
    bound = set(re.findall(r"(?:function|subroutine)\s+(i3rc_\w+)\s*\(", s))
    assert {"i3rc_new_Integrator", "i3rc_set_phase_table", "i3rc_specifyParameters", "i3rc_computeRadiativeTransfer",
            "i3rc_reportResults", "i3rc_copy_Integrator", "i3rc_finalize_Integrator", "i3rc_isReady_Integrator",
            "i3rc_last_message"} <= bound
    for sym in bound:
        assert re.search(rf"\b{sym}\s*\(", HEADER), f"{sym} is not declared in include/i3rc_b200.h"


def _c_struct_fields(name):
    body = re.search(r"typedef struct\s*\{([^{}]*)\}\s*" + name + r"\s*;", HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
        out += [n.strip().lstrip("*").strip() for n in names.split(",")]
    return out


def _f_type_fields(src, name):
    body = re.search(r"type,\s*bind\(c\)(?:,\s*public)?\s*::\s*" + name + r"\s*\n(.*?)end type", src, flags=re.S | re.I).group(1)
    body = body.replace("&\n", " ")
    out = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" in line:
            out += [re.sub(r"=.*", "", n).strip() for n in re.split(r",(?![^()]*\))", line.split("::", 1)[1])]
    return [n for n in out if n]


def test_shim_struct_mirrors_match_the_header_field_by_field():
    main, ill = _src("monteCarloRadiativeTransfer_b200.f95"), _src("monteCarloIllumination_b200.f95")
    assert _f_type_fields(ill, "i3rc_photon_source") == _c_struct_fields("i3rc_photon_source")
    assert _f_type_fields(main, "i3rc_phase_table") == _c_struct_fields("i3rc_phase_table")
    assert _f_type_fields(main, "i3rc_params") == _c_struct_fields("i3rc_params")
    # the source kinds and the presence bits the shim hard-codes
    kinds = dict(re.findall(r"(I3RC_SRC_\w+)\s*=\s*(\d+)", ill))
    for k, v in kinds.items():
        assert re.search(rf"{k}\s*=\s*{v}\b", HEADER), k
    bits = dict(re.findall(r"I3RC_P_(\w+)\s*=\s*1u\s*<<\s*(\d+)", HEADER))
    for name, power in re.findall(r"present\((\w+)\)\)\s*then;?\s*(?:\n\s*)?p%present = ior\(p%present, 2\*\*(\d+)\)", main):
        assert bits[name] == power, name


def test_mersenne_twister_of_the_replacement_module():
    """A statement-for-statement Python transliteration of fortran/RandomNumbersForMC_b200.f95 (seedState,
    initialize_vector, nextState, getRandomInt) against the published mt19937ar known answers and numpy's MT19937."""
    mask, N, M = 0xFFFFFFFF, 624, 397

    def seed_state(s):
        st, prev = [0] * N, s & mask
        st[0] = prev
        for i in range(1, N):
            prev = (1812433253 * (prev ^ (prev >> 30)) + i) & mask
            st[i] = prev
        return st

    def initialize_vector(seed):
        st, i, j = seed_state(19650218), 1, 0
        for _ in range(max(N, len(seed)), 0, -1):
            prev, cur = st[i - 1], st[i]
            st[i] = ((cur ^ (((prev ^ (prev >> 30)) * 1664525) & mask)) + (seed[j] & mask) + j) & mask
            i, j = i + 1, j + 1
            if i >= N:
                st[0], i = st[N - 1], 1
            if j >= len(seed):
                j = 0
        for _ in range(N - 1, 0, -1):
            prev, cur = st[i - 1], st[i]
            st[i] = ((cur ^ (((prev ^ (prev >> 30)) * 1566083941) & mask)) - i + 4294967296) & mask
            i += 1
            if i >= N:
                st[0], i = st[N - 1], 1
        st[0] = 0x80000000
        return st

    def draws(st, n):
        out, cur = [], N
        for _ in range(n):
            if cur >= N:
                for k in range(N):
                    y = (st[k] & 0x80000000) | (st[(k + 1) % N] & 0x7FFFFFFF)
                    y = (y >> 1) ^ st[(k + M) % N]
                    if st[(k + 1) % N] & 1:
                        y ^= 0x9908B0DF
                    st[k] = y
                cur = 0
            y = st[cur]
            cur += 1
            y ^= y >> 11
            y ^= (y << 7) & 0x9D2C5680
            y ^= (y << 15) & 0xEFC60000
            y ^= y >> 18
            out.append(y & mask)
        return out

    assert draws(initialize_vector([0x123, 0x234, 0x345, 0x456]), 5) == [1067595299, 955945823, 477289528, 4107218783, 4228976476]
    bg = np.random.MT19937()
    bg._legacy_seeding(np.array([10, 1], dtype=np.uint32))  # (/ iseed, batch /)
    assert draws(initialize_vector([10, 1]), 1500) == bg.random_raw(1500).tolist()


@pytest.fixture(scope="module")
def replay():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "ctest"), "-s"])
    return os.path.join(ROOT, "tests", "ctest", "_build", "shim_replay")


def test_c_replay_links_and_refuses_to_run_without_a_gpu(replay, cuda_absent):
    r = subprocess.run([replay], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_replay_of_the_shim_reproduces_the_python_mirror(replay, cuda):
    """Same problem, same seeds (/ iseed, batch /), same call order: the C program and the Python mirror trace the same
    photons (Philox streams), so their batch results agree to float32 summation order."""
    from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, new_Integrator_dense,
                                                                        reportResults, specifyParameters)
    from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
    from i3rc_monte_carlo_model_b200.scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable
    r = subprocess.run([replay, "4", "20000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(t) for t in ln.split()] for ln in r.stdout.splitlines()])
    assert rows.shape == (4, 6)
    f32 = np.float32
    x, y, z = f32(125.0) * np.arange(5, dtype=f32), f32(250.0) * np.arange(3, dtype=f32), f32(100.0) * np.arange(4, dtype=f32)
    ext = np.zeros((4, 2, 3), f32)
    for k in range(3):
        ext[:2, :, k] = f32(0.004) * (f32(1.0) + f32(0.25) * f32(k))
        ext[2:, :, k] = f32(0.012) * (f32(1.0) + f32(0.25) * f32(k))
    table = new_PhaseFunctionTable([new_PhaseFunction((f32(0.85) ** np.arange(1, 17)).astype(f32))], [1.0])
    I = new_Integrator_dense(x, y, z, ext, np.ones((4, 2, 3, 1), f32), np.full((4, 2, 3, 1), 0.95, f32), np.ones((4, 2, 3, 1), np.int32),
                             [table], backend=cuda)
    specifyParameters(I, surfaceAlbedo=0.2, minInverseTableSize=10001)
    specifyParameters(I, minForwardTableSize=10001, intensityMus=[1.0], intensityPhis=[0.0], computeIntensity=True)
    specifyParameters(I, useRayTracing=True, useRussianRoulette=True, useRussianRouletteForIntensity=True, zetaMin=0.3,
                      useHybridPhaseFunsForIntenCalcs=False, hybridPhaseFunWidth=0.0, numOrdersOrigPhaseFunIntenCalcs=0,
                      limitIntensityContributions=False, maxIntensityContribution=0.0)
    for b in range(1, 5):
        computeRadiativeTransfer(I, new_RandomNumberSequence([10, b]), new_PhotonStream(0.5, 0.0, numberOfPhotons=20000))
        res = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity", "fluxUp")
        want = [b, res["meanFluxUp"], res["meanFluxDown"], res["meanFluxAbsorbed"], res["meanIntensity"][0], np.mean(res["fluxUp"])]
        assert np.allclose(rows[b - 1], want, rtol=3e-5, atol=2e-6), (b, rows[b - 1], want)
    assert abs(rows[:, 1].mean() + rows[:, 3].mean() + 0.8 * rows[:, 2].mean() - 1.0) < 5e-3  # energy closure
