"""python -m cuda_dct_idct_b200 <input_image> <output_image> [retained_coefficients]

The reference programs' command line (main_newAppr.cu:28-31): load a grayscale image, run
DCT -> quantise -> IDCT on the GPU, save the reconstruction at JPEG quality 100."""
import sys

from . import api, imageio


def main(argv):
    if len(argv) not in (3, 4):
        print(f"Usage: {argv[0]} <input_image> <output_image> [retained_coefficients 1..64]", file=sys.stderr)
        return 1
    plan = api.Plan(keep=api.zigzag_mask(int(argv[3]))) if len(argv) == 4 else None
    mse, peen = imageio.transform_file(argv[1], argv[2], plan=plan)
    print(f"Image saved successfully to {argv[2]}  (MSE {mse:.4f}, PEEN {peen:.4f} %)")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
