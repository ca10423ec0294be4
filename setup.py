from setuptools import find_packages, setup

setup(
    name="transcriptclean_b200",
    version="0.1.0",
    description="B200-native correction path of TranscriptClean (CUDA, sm_100a) behind the transcriptclean CLI",
    packages=find_packages(include=["transcriptclean_b200*"]),
    package_data={"transcriptclean_b200": ["csrc/*.so", "csrc/*.cu", "csrc/*.cuh"]},
    entry_points={"console_scripts": ["transcriptclean=transcriptclean_b200.cli:main"]},
)
