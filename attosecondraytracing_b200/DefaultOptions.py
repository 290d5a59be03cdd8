"""Default option dictionaries of the reference's driver (ART/DefaultOptions.py): the keys and default values a
config script may leave out.  The `plot_*` / render options are accepted for compatibility; drawing is out of
scope of this package (ModuleAnalysisAndPlots offers the figures' data instead)."""

DefaultAnalysisOptions = {
    "verbose": True,
    "plot_Render": False, "maxRaysToRender": 200, "OEPointsToRender": 3000, "OEPointsScale": 5,
    "draw_mesh": False, "cycle_ray_colors": False, "DrawAiryAndFourier": True,
    "plot_SpotDiagram": False, "plot_DelaySpotDiagram": False, "plot_IntensitySpotDiagram": False,
    "plot_IncidenceSpotDiagram": False, "plot_DelayGraph": False, "plot_IntensityGraph": False,
    "plot_IncidenceGraph": False, "plot_DelayMirrorProjection": False, "plot_IntensityMirrorProjection": False,
    "plot_IncidenceMirrorProjection": False,
    "save_results": True,
}

DefaultSourceProperties = {
    "Divergence": 0,        # half-angle in rad
    "SourceSize": 0,        # diameter in mm
    "Wavelength": 50e-6,    # mm
    "DeltaFT": 1,           # fs
    "NumberRays": 1000,
}

DefaultDetectorOptions = {
    "ReflectionNumber": -1,         # analyse the bundle after the last optical element
    "ManualDetector": False,
    "DetectorCentre": None,
    "DetectorNormal": None,
    "DistanceDetector": None,
    "AutoDetectorDistance": False,  # search the optimal detector distance first
    "OptFor": "intensity",
}
