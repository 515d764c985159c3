"""solve(batch) adapters for sbdart_b200.frontend.Sbdart.run: CPU oracle and CUDA."""
import numpy as np

from oracle import oracle


def _set_oracle_surface(b, i):
    spec = b["surface"]
    st = spec.states[int(b["surf"][i])]
    oracle.set_bdref(spec.model.ibdrf, spec.model.params(), *((st["nr"], st["ni"], st["rsw"]) if st else (0, 0, 0)))


def solve_oracle(b, nthreads=8):
    bins = b["bins"]
    if "umu" not in b and "surface" in b:          # BRDF flux run: bin by bin (the model is file-scope state)
        outs = {k: [] for k in ("rfldir", "rfldn", "flup", "dfdt", "uavg", "status")}
        for i in range(len(bins)):
            _set_oracle_surface(b, i)
            r = oracle.disort(
                b["dtauc"][i], b["ssalb"][i], b["pmom"][i], nstr=b["nstr"], temper=b["temper"][0],
                fbeam=bins["fbeam"][i], umu0=bins["umu0"][i], fisot=bins["fisot"][i], btemp=bins["btemp"][i],
                ttemp=bins["ttemp"][i], temis=bins["temis"][i], wvnmlo=bins["wvnmlo"][i],
                wvnmhi=bins["wvnmhi"][i], plank=bool(bins["plank"][i]), lamber=False)
            for k in outs:
                outs[k].append(r[k])
        return {k: np.array(v) for k, v in outs.items()}
    if "umu" not in b:
        return oracle.disort_flux_batch(
            b["dtauc"], b["ssalb"], b["pmom"], nstr=b["nstr"], fbeam=bins["fbeam"], umu0=bins["umu0"],
            albedo=bins["albedo"], plank=bins["plank"], wvnmlo=bins["wvnmlo"], wvnmhi=bins["wvnmhi"],
            btemp=bins["btemp"], ttemp=bins["ttemp"], temis=bins["temis"], fisot=bins["fisot"],
            temper=b["temper"], col=bins["col"], nthreads=nthreads)
    outs = {k: [] for k in ("rfldir", "rfldn", "flup", "uu", "status")}
    for i in range(len(bins)):
        if "surface" in b:
            _set_oracle_surface(b, i)
        r = oracle.disort(
            b["dtauc"][i], b["ssalb"][i], b["pmom"][i], nstr=b["nstr"], temper=b["temper"][0],
            lamber="surface" not in b,
            umu=b["umu"], phi=b["phi"], fbeam=bins["fbeam"][i], umu0=bins["umu0"][i],
            phi0=bins["phi0"][i], fisot=bins["fisot"][i], albedo=bins["albedo"][i],
            btemp=bins["btemp"][i], ttemp=bins["ttemp"][i], temis=bins["temis"][i],
            wvnmlo=bins["wvnmlo"][i], wvnmhi=bins["wvnmhi"][i], plank=bool(bins["plank"][i]),
            onlyfl=False, corint=b.get("corint", False))
        for k in outs:
            outs[k].append(r[k])
    return {k: np.array(v) for k, v in outs.items()}


def make_solve_cuda(solver, packed=False):
    """packed: with a level selection, uu comes back as [B][nphi][nsel][numu] (+ "uu_levels")."""
    def solve(b):
        bins = b["bins"]
        if "surface" in b:
            import sbdart_b200 as sb
            tab = b["surface"].tables(b["nstr"], sb.quadrature)
            solver.set_surfaces(b["nstr"], tab["bdr"], tab["bem"], tab.get("rmu"), tab.get("emu"))
            bins = bins.copy()
            bins["albedo"] = sb.surface_albedo(b["surf"])
        try:
            return solver.disort_batch(b["dtauc"], b["ssalb"], b["pmom"], bins, nstr=b["nstr"],
                                       temper=b["temper"], umu=b.get("umu"), phi=b.get("phi"), uu_levels=b.get("uu_levels"),
                                       corint=b.get("corint", False), uu_packed=packed)
        finally:
            if "surface" in b:
                solver.set_surfaces()
    return solve
