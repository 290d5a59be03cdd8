"""B200-native ray-bundle engine behind ART's RayTracingCalculation / OpticalChain.get_output_rays."""
__version__ = "0.1.0"
