timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
cd refrakt_b200 && time ./rfk_render --genome ../tests/fixtures/electricsheep.247.11256.flam3 --variations ../tests/fixtures/variations.yaml --width 7680 --height 4320 --supersample 2 --quality 100 --out ../gpurun_out/frame_8k_ss2.png 2>&1 | tail -8; ls -la ../gpurun_out/frame_8k_ss2.png
