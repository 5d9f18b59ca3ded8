'''
drone3d/utils/load_utils.py of the reference.  The reference ships its assets (a mesh, two CSV files of an external
solver, textures) inside the package; they are data of the reference and are not copied here.  Files are looked up in
this package's `assets/` folder, then in $RACELINE_ASSETS, then in a reference checkout at /root/reference.
'''
import os


def get_assets_folder() -> str:
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'assets')
    for folder in (here, os.environ.get('RACELINE_ASSETS', ''), '/root/reference/drone3d/assets'):
        if folder and os.path.isdir(folder) and len(os.listdir(folder)) > 0:
            return folder
    return here


def get_assets_file(file: str) -> str:
    for folder in (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'assets'),
                   os.environ.get('RACELINE_ASSETS', ''), '/root/reference/drone3d/assets'):
        if folder and os.path.exists(os.path.join(folder, file)):
            return os.path.join(folder, file)
    return os.path.join(get_assets_folder(), file)
