"""The reference's driver (ARTmain.py:99-345) without its plotting: complete the option dictionaries, trace the
optical chain(s), set up / optimise the detector, summarise -- the caller of the hot path this package
accelerates.  Same function names, arguments and return values; every ray-level step runs on the GPU
(OpticalChain.get_output_rays, Detector.autoplace, FindOptimalDistance, GetResultSummary).

Differences a config script will notice: nothing is drawn (options `plot_*` only produce a notice; use
ModuleAnalysisAndPlots.SpotDiagramData / DelayGraphData / MirrorProjectionData for the figures' data); the
detector-distance optimisation uses ALL rays in closed form instead of 1000 randomly drawn ones
(ARTmain.py:168-171), so it is deterministic."""
from __future__ import annotations

import copy

from . import ModuleAnalysisAndPlots as mplots
from . import ModuleDetector as mdet
from . import ModuleOpticalChain as moc
from . import ModuleProcessing as mp


def complete_defaults(SourceProperties, DetectorOptions, AnalysisOptions):
    """The three option dictionaries completed with the defaults (ARTmain.py:99-110).  The defaults
    themselves are not modified (the reference updates its module-level dictionaries in place)."""
    from .DefaultOptions import DefaultAnalysisOptions, DefaultDetectorOptions, DefaultSourceProperties
    out = []
    for default, given in ((DefaultSourceProperties, SourceProperties), (DefaultDetectorOptions, DetectorOptions),
                           (DefaultAnalysisOptions, AnalysisOptions)):
        d = copy.deepcopy(default)
        d.update(given)
        out.append(d)
    return tuple(out)


def setup_detector(OpticalChain, DetectorOptions, RayList=None):
    """The Detector the options describe: placed manually (centre + normal given) or automatically at
    DistanceDetector along the central ray of RayList (ARTmain.py:113-144)."""
    ref_point = OpticalChain.optical_elements[DetectorOptions["ReflectionNumber"]].position
    if DetectorOptions["ManualDetector"]:
        for key in ("DetectorCentre", "DetectorNormal"):
            if DetectorOptions[key] is None:
                raise RuntimeError(f'For manual detector placement you need to specify "{key}" in the '
                                   '"DetectorOptions"-dictionary.')
        return mdet.Detector(ref_point, DetectorOptions["DetectorCentre"], DetectorOptions["DetectorNormal"])
    if DetectorOptions["DistanceDetector"] is None:
        raise RuntimeError('For automatic detector placement you need to specify "DistanceDetector" in the '
                           '"DetectorOptions"-dictionary.')
    if RayList is None:
        raise RuntimeError('For automatic detector placement you need to add a RayList as an input (selected from '
                           'the "RayListHistory" by the index DetectorOptions["ReflectionNumber"]).')
    detector = mdet.Detector(ref_point)
    detector.autoplace(RayList, DetectorOptions["DistanceDetector"])
    return detector


def optimize_detector(RayListAnalysed, Detector, DetectorOptions, verbose=True, maxRaystoConsider=None,
                      IntensityWeighted=False, Amplitude=None, Precision=3):
    """Detector moved to the distance that optimises DetectorOptions["OptFor"] ("spotsize", "duration" or
    "intensity"); returns (detector, spot size SD in mm, duration SD in fs) (ARTmain.py:147-190).
    maxRaystoConsider is accepted and ignored: all rays enter the closed-form search."""
    opt_for = DetectorOptions["OptFor"]
    detector, spot, duration = mp.FindOptimalDistance(Detector, RayListAnalysed, opt_for, Amplitude, Precision,
                                                      IntensityWeighted, verbose)
    if verbose:
        text = f"The optimal detector distance is {detector.get_distance():.3f} mm, with"
        if IntensityWeighted:
            text += " intensity-weighted"
        if opt_for in ("intensity", "spotsize"):
            text += f" spatial std of {spot * 1e3:.3g} μm"
        if opt_for in ("intensity", "duration"):
            text += f" temporal std of {duration:.3g} fs."
        print(text, flush=True)
    return detector, spot, duration


def run_ART(OpticalChain, SourceProperties, DetectorOptions, AnalysisOptions, loop=False):
    """Trace one optical chain and analyse the bundle after element DetectorOptions["ReflectionNumber"]
    (ARTmain.py:248-300).  Returns (OpticalChain, Detector, ETransmission %, SpotSizeSD mm, DurationSD fs)."""
    rays = OpticalChain.get_output_rays()[DetectorOptions["ReflectionNumber"]]
    transmission = mplots.getETransmission(OpticalChain.source_rays, rays)
    verbose = AnalysisOptions["verbose"]
    if verbose:
        print("_" * 99, flush=True)
        if isinstance(OpticalChain.description, str) and OpticalChain.description:
            print("***" + OpticalChain.description + "*** :")
        if OpticalChain.loop_variable_name is not None and OpticalChain.loop_variable_value is not None:
            print(f"For {OpticalChain.loop_variable_name} = {OpticalChain.loop_variable_value:f}:\n")
            print(f"The optical setup has an energy transmission of {transmission:.1f}%.\n")
    detector = setup_detector(OpticalChain, DetectorOptions, rays)
    if DetectorOptions["AutoDetectorDistance"]:
        detector, spot, duration = optimize_detector(rays, detector, DetectorOptions, verbose, IntensityWeighted=True)
    else:
        spot, duration = mplots.GetResultSummary(detector, rays, verbose)
    if verbose:
        print("_" * 99 + "\n")
    if not loop and any(v for k, v in AnalysisOptions.items() if k.startswith("plot_")):
        print("[attosecondraytracing_b200] plotting is out of scope of this package: use "
              "ModuleAnalysisAndPlots.SpotDiagramData / DelayGraphData / MirrorProjectionData", flush=True)
    return OpticalChain, detector, transmission, spot, duration


def main(OpticalChainList, SourceProperties, DetectorOptions, AnalysisOptions, save_file_name=None):
    """run_ART for an OpticalChain or a list of them (e.g. from OpticalChain.get_OE_loop_list); returns the
    dictionary of lists {"OpticalChain", "Detector", "ETransmission", "SpotSizeSD", "DurationSD"} and saves it
    with save_compressed when AnalysisOptions["save_results"] (ARTmain.py:304-345)."""
    SourceProperties, DetectorOptions, AnalysisOptions = complete_defaults(SourceProperties, DetectorOptions,
                                                                          AnalysisOptions)
    names = ["OpticalChain", "Detector", "ETransmission", "SpotSizeSD", "DurationSD"]
    kept = {name: [] for name in names}
    if isinstance(OpticalChainList, moc.OpticalChain):
        chains, loop = [OpticalChainList], False
    elif isinstance(OpticalChainList, list):
        chains, loop = OpticalChainList, True
    else:
        raise ValueError("The supplied OpticalChain is neither an OpticalChain-object, nor a list of those, as it "
                         "should be.")
    for i, chain in enumerate(chains):
        print(f"Optical Chain {i}/{len(chains)} ", end="", flush=True)
        for name, value in zip(names, run_ART(chain, SourceProperties, DetectorOptions, AnalysisOptions, loop)):
            kept[name].append(value)
    if AnalysisOptions["save_results"]:
        mp.save_compressed(kept, save_file_name)
    return kept
