"""Population parameter files: the reference's per-population XML, read and written.

Format (`tutorial_data/xmldat/tut_EnvironAlt.xml`, parsed in the reference by
`io/ParamProvider2.cpp` + `io/qhgXML.cpp`):

    <class name="..." species_name="..." species_id="...">
      <module name="ATanDeath" [id="Alt"]> <param name="ATanDeath_max_age" value="60.0"/> ... </module>
      <priorities> <prio name="GetOld" value="1"/> ... </priorities>
    </class>

Action names are `Name` or `Name[id]` (`actions/Action.cpp:11-19`); they key both the
`<prio>` table and the parameter groups.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field


def base_is_multi(name: str) -> bool:
    return name.startswith("MultiEvaluator")


@dataclass
class PopParams:
    class_name: str
    species_name: str = "sapiens"
    species_id: int = 104
    modules: dict = field(default_factory=dict)   # action name (with [id]) -> {param: value string}
    prios: dict = field(default_factory=dict)     # action name -> int

    def to_xml(self) -> str:
        out = [f'<class name="{self.class_name}" species_name="{self.species_name}" species_id="{self.species_id}">']
        for name, pars in self.modules.items():
            if "[" in name:
                base, ident = name[:-1].split("[", 1)
                out.append(f'  <module name="{base}" id="{ident}">')
            else:
                out.append(f'  <module name="{name}">')
            sub = {"AltPref": "Alt", "AltCapPref": "Alt", "NPPPref": "NPP"}  # poly-lines of the evaluators inside a MultiEvaluator
            for k, v in pars.items():
                if base_is_multi(name) and k in sub:
                    continue
                out.append(f'    <param name="{k}" value="{v}"/>')
            if base_is_multi(name):  # tutorial_data/xmldat/tut_EnvironCapAlt.xml: the evaluators are sub-modules
                for ident in ("Alt", "NPP"):
                    out.append(f'    <module name="SingleEvaluator" id="{ident}">')
                    for k, v in pars.items():
                        if sub.get(k) == ident:
                            out.append(f'      <param name="{k}" value="{v}"/>')
                    out.append("    </module>")
            out.append("  </module>")
        out.append("  <priorities>")
        for name, p in self.prios.items():
            out.append(f'    <prio name="{name}" value="{p}"/>')
        out.append("  </priorities>")
        out.append("</class>")
        return "\n".join(out) + "\n"

    def write(self, path: str) -> str:
        with open(path, "w") as f:
            f.write(self.to_xml())
        return path

    @staticmethod
    def from_xml(text: str, class_name: str | None = None) -> "PopParams":
        root = ET.fromstring(text)
        classes = [root] if root.tag == "class" else list(root.iter("class"))
        for c in classes:
            if class_name is None or c.get("name") == class_name:
                pp = PopParams(c.get("name"), c.get("species_name", ""), int(c.get("species_id", "0")))
                for m in c.findall("module"):
                    name = m.get("name")
                    if m.get("id"):
                        name = f'{name}[{m.get("id")}]'
                    pp.modules[name] = {p.get("name"): p.get("value") for p in m.iter("param")}
                pr = c.find("priorities")
                if pr is not None:
                    for p in pr.findall("prio"):
                        pp.prios[p.get("name")] = int(p.get("value"))
                return pp
        raise KeyError(class_name)

    def copy(self) -> "PopParams":
        return PopParams(self.class_name, self.species_name, self.species_id,
                         {k: dict(v) for k, v in self.modules.items()}, dict(self.prios))


def tut_environ_alt(K: float = 20.0) -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_EnvironAlt.xml` (values restated, `Verhulst_K` adjustable)."""
    return PopParams(
        "tut_EnvironAltPop",
        modules={
            "ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"},
            "WeightedMove": {"WeightedMove_prob": "0.07"},
            "Fertility": {"Fertility_interbirth": "2.0", "Fertility_max_age": "50.0", "Fertility_min_age": "15.0"},
            "Verhulst": {"Verhulst_b0": "0.8", "Verhulst_d0": "0.001", "Verhulst_theta": "0.1", "Verhulst_K": repr(float(K))},
            "SingleEvaluator[Alt]": {"AltCapPref": "-0.1 0 0.1 0.01 1500 1.0 2000 1 3000 -9999"},
        },
        prios={"GetOld": 1, "ATanDeath": 2, "WeightedMove": 3, "SingleEvaluator[Alt]": 4,
               "Fertility": 5, "RandomPair": 6, "Verhulst": 7},
    )


def tut_environ_cap_alt() -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_EnvironCapAlt.xml` (values restated)."""
    return PopParams(
        "tut_EnvironCapAltPop",
        modules={
            "ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"},
            "WeightedMove": {"WeightedMove_prob": "0.07"},
            "Fertility": {"Fertility_interbirth": "2.0", "Fertility_max_age": "50.0", "Fertility_min_age": "15.0"},
            "NPPCapacity": {"NPPCap_K_max": "38.4716796875", "NPPCap_K_min": "0.0", "NPPCap_NPP_max": "1.0576171875",
                            "NPPCap_NPP_min": "0.0", "NPPCap_coastal_factor": "0.4", "NPPCap_coastal_max_latitude": "66.0",
                            "NPPCap_coastal_min_latitude": "50.0", "NPPCap_water_factor": "0.5349609375", "NPPCap_efficiency": "1"},
            "MultiEvaluator[NPP+Alt]": {"Multi_weight_alt": "0.2", "Multi_weight_npp": "0.8",
                                        "AltPref": "-0.1 0 0.1 0.01 1500 1.0 2000 1 3000 -9999"},
            "VerhulstVarK": {"Verhulst_b0": "0.8", "Verhulst_d0": "0.001", "Verhulst_theta": "0.1"},
        },
        prios={"NPPCapacity": 1, "GetOld": 2, "ATanDeath": 3, "WeightedMove": 4, "MultiEvaluator[NPP+Alt]": 5,
               "Fertility": 6, "RandomPair": 7, "VerhulstVarK": 8},
    )


def tut_sexual(K: float = 30.0, move_prob: float = 0.05) -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_Sexual.xml` (values restated): RandomMove instead of WeightedMove, and
    the ageing / death actions AFTER pairing, births and the move."""
    return PopParams(
        "tut_SexualPop",
        modules={
            "ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"},
            "RandomMove": {"RandomMove_prob": repr(float(move_prob))},
            "Fertility": {"Fertility_interbirth": "2.0", "Fertility_max_age": "50.0", "Fertility_min_age": "14.0"},
            "Verhulst": {"Verhulst_b0": "0.8", "Verhulst_d0": "0.001", "Verhulst_theta": "0.1", "Verhulst_K": repr(float(K))},
        },
        prios={"GetOld": 8, "ATanDeath": 10, "RandomMove": 7, "Fertility": 2, "RandomPair": 3, "Verhulst": 6},
    )


def tut_move(move_prob: float = 0.1) -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_Move.xml` (values restated): ageing, ATanDeath, RandomMove; no births."""
    return PopParams(
        "tut_MovePop",
        modules={
            "ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"},
            "RandomMove": {"RandomMove_prob": repr(float(move_prob))},
        },
        prios={"GetOld": 8, "ATanDeath": 10, "RandomMove": 9},
    )


def tut_old_age_die() -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_OldAgeDie.xml` (values restated): ageing and ATanDeath only."""
    return PopParams(
        "tut_OldAgeDiePop",
        modules={"ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"}},
        prios={"GetOld": 8, "ATanDeath": 10},
    )


def tut_partheno(K: float = 20.0, move_prob: float = 0.01) -> PopParams:
    """Parameter set of `tutorial_data/xmldat/tut_Partheno.xml` (values restated): females only, no pairing action --
    every agent counts as mated (populations/tut_ParthenoPop.cpp:107-119), all newborns are female."""
    return PopParams(
        "tut_ParthenoPop",
        modules={
            "ATanDeath": {"ATanDeath_max_age": "60.0", "ATanDeath_range": "6.0", "ATanDeath_slope": "1.0"},
            "RandomMove": {"RandomMove_prob": repr(float(move_prob))},
            "Fertility": {"Fertility_interbirth": "2.0", "Fertility_max_age": "50.0", "Fertility_min_age": "15.0"},
            "Verhulst": {"Verhulst_b0": "0.8", "Verhulst_d0": "0.001", "Verhulst_theta": "0.1", "Verhulst_K": repr(float(K))},
        },
        prios={"GetOld": 8, "ATanDeath": 10, "RandomMove": 7, "Fertility": 2, "Verhulst": 6},
    )


def tut_static() -> PopParams:
    """`tutorial_data/xmldat/tut_Static.xml`: a population without actions (populations/tut_StaticPop.cpp:16-21)."""
    return PopParams("tut_StaticPop")


def tut_environ_alt_confined(K: float, x: float, y: float, r_km: float, prio: int = 9) -> PopParams:
    """tut_EnvironAltPop's parameter set plus ConfinedMove (actions/ConfinedMove.cpp; attributes ConfinedMove_x/_y in degrees,
    ConfinedMove_r in km).  No shipped population of this size carries it (21 of the OoA* classes do): the class name is the
    probe class `tut_EnvironAltConfPop` of oracle/ref_driver.cpp, the oracle and the CUDA library."""
    par = tut_environ_alt(K)
    par.class_name = "tut_EnvironAltConfPop"
    par.modules["ConfinedMove"] = {"ConfinedMove_x": repr(float(x)), "ConfinedMove_y": repr(float(y)), "ConfinedMove_r": repr(float(r_km))}
    par.prios["ConfinedMove"] = prio
    return par


def tut_environ_alt_variants(K: float, move_rand: bool = True, sig_death: bool = True, move_prob: float = 0.2) -> PopParams:
    """tut_EnvironAltPop's parameter set with WeightedMoveRand (actions/WeightedMoveRand.cpp) in place of WeightedMove and / or
    SigDeath (actions/SigDeath.cpp) in place of ATanDeath: the probe class `tut_EnvironAltVarPop` (all four actions are
    registered; the <prio> entries decide which run)."""
    par = tut_environ_alt(K)
    par.class_name = "tut_EnvironAltVarPop"
    par.modules["WeightedMoveRand"] = {"WeightedMoveRand_prob": repr(float(move_prob))}
    par.modules["SigDeath"] = {"SigDeath_max_age": "60.0", "SigDeath_range": "6.0", "SigDeath_slope": "0.5"}
    if move_rand:
        par.prios["WeightedMoveRand"] = par.prios.pop("WeightedMove")
    if sig_death:
        par.prios["SigDeath"] = par.prios.pop("ATanDeath")
    return par


def tut_environ_alt_ext(K: float, cond_mode: int = -1, perm_pair: bool = False, move_stats: int = -1, move_prob: float = 0.2) -> PopParams:
    """tut_EnvironAltPop's parameter set for the probe classes `tut_EnvironAltCond<m>Pop` (ExtProbePop<m> of oracle/ref_driver.cpp):
    CondWeightedMove (actions/CondWeightedMove.cpp, a SimpleCondition of mode m over the altitudes) in place of WeightedMove when
    cond_mode >= 0, RandPermPair (actions/RandPermPair.cpp) in place of RandomPair, and MoveStats (actions/MoveStats.cpp) with
    MoveStats_Mode = move_stats (0 first, 1 minimum, 2 last) as the step's last action when move_stats >= 0."""
    par = tut_environ_alt(K)
    par.class_name = f"tut_EnvironAltCond{max(cond_mode, 0)}Pop"
    par.modules["CondWeightedMove"] = {"CondWeightedMove_prob": repr(float(move_prob))}
    par.modules["MoveStats"] = {"MoveStats_Mode": str(int(max(move_stats, 0)))}
    if cond_mode >= 0:
        par.prios["CondWeightedMove"] = par.prios.pop("WeightedMove")
    if perm_pair:
        par.prios["RandPermPair"] = par.prios.pop("RandomPair")
    if move_stats >= 0:
        par.prios["MoveStats"] = max(par.prios.values()) + 1
    return par


def tut_environ_alt_genetic(K: float, genome_size: int, num_crossover: int, mutation_rate: float, bits_per_nuc: int = 1) -> PopParams:
    """tut_EnvironAltPop's parameter set plus Genetics (actions/Genetics.cpp) with 1-bit (genes/BitGeneUtils.cpp) or 2-bit
    (genes/GeneUtils.cpp) nucleotides: the probe classes `tut_EnvironAltGenPop` / `tut_EnvironAltGen2bitPop` of
    oracle/ref_driver.cpp, the oracle and the CUDA library, which pin the Genetics action against the reference's own."""
    par = tut_environ_alt(K)
    par.class_name = "tut_EnvironAltGen2bitPop" if bits_per_nuc == 2 else "tut_EnvironAltGenPop"
    par.modules["Genetics"] = {"Genetics_genome_size": str(int(genome_size)), "Genetics_num_crossover": str(int(num_crossover)),
                               "Genetics_mutation_rate": repr(float(mutation_rate)), "Genetics_create_new_genome": "0",
                               "Genetics_bits_per_nuc": str(int(bits_per_nuc)), "Genetics_initial_muts": "none"}
    par.prios["Genetics"] = 9
    return par


def ooa_nav_gen(genome_size: int = 4096, num_crossover: int = -1, mutation_rate: float = 1e-5) -> PopParams:
    """`OoANavGenPop` (populations/OoANavGenPop.cpp:33-97) without Navigate: the genetic population of config C3.
    The reference ships no parameter file for it; ecological values follow tut_EnvironCapAlt.xml, the Genetics values are
    the ones SURVEY.md §8d declares (4096 one-bit sites, free recombination, mutation rate 1e-5)."""
    base = tut_environ_cap_alt()
    mods = {k: dict(v) for k, v in base.modules.items() if k not in ("ATanDeath", "MultiEvaluator[NPP+Alt]")}
    mods["OldAgeDeath"] = {"OAD_max_age": "60.0", "OAD_uncertainty": "0.1"}
    mods["MultiEvaluator[Alt+NPP]"] = {"Multi_weight_alt": "0.2", "Multi_weight_npp": "0.8",
                                       "AltCapPref": "-0.1 0 0.1 0.01 1500 1.0 2000 1 3000 -9999",
                                       "NPPPref": "0 0 5 0.2 38.5 1.0"}
    mods["Genetics"] = {"Genetics_genome_size": str(int(genome_size)), "Genetics_num_crossover": str(int(num_crossover)),
                        "Genetics_mutation_rate": repr(float(mutation_rate)), "Genetics_create_new_genome": "0",
                        "Genetics_bits_per_nuc": "1", "Genetics_initial_muts": "none"}
    # every registered action needs its module entry (core/Prioritizer.cpp getActionParams), with or without a <prio>: Navigate's
    # values follow useful_stuff/realistic_sap.xml; config C5 adds the priority
    mods["Navigate"] = {"Navigate_decay": "-0.001", "Navigate_dist0": "150.0", "Navigate_prob0": "0.1", "Navigate_min_dens": "0.0",
                        "Navigate_bridge_prob": "0.3"}
    return PopParams("OoANavGenPop", modules=mods,
                     prios={"NPPCapacity": 1, "GetOld": 2, "OldAgeDeath": 3, "WeightedMove": 4, "MultiEvaluator[Alt+NPP]": 5,
                            "Fertility": 6, "RandomPair": 7, "VerhulstVarK": 8, "Genetics": 9})


# default WELL512 state of the reference (app/SimParams.cpp:82-87): 16 words of seed material
DEFAULT_STATE = (
    0x2ef76080, 0x1bf121c5, 0xb222a768, 0x6c5d388b, 0xab99166e, 0x326c9f12, 0x3354197a, 0x7036b9a5,
    0xb08c9e58, 0x3362d8d3, 0x037e5e95, 0x47a1ff2f, 0x740ebb34, 0xbf27ef0d, 0x70055204, 0xd24daa9a,
)


def seed_state(seed: int):
    """16-word RNG state for an integer seed (seed 0 = the reference's default table)."""
    import numpy as np
    st = np.array(DEFAULT_STATE, dtype=np.uint32)
    if seed:
        rng = np.random.default_rng(int(seed))
        st = st ^ rng.integers(0, 2 ** 32, size=16, dtype=np.uint64).astype(np.uint32)
    return st
