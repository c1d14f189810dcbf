"""Same attribute surface as the reference's dvmvs/config.py:4-51 (class attributes, edited in source or
overridden by the caller).  Model constructors read train_n_depth_levels / train_min_depth / train_max_depth
(reference fusionnet/model.py:170,231-232)."""
import os
import time

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Config:
    # training settings (kept for API parity; training is outside this package's scope)
    train_image_width = 256
    train_image_height = 256
    train_min_depth = 0.25
    train_max_depth = 20.0
    train_n_depth_levels = 64
    train_minimum_pose_distance = 0.125
    train_maximum_pose_distance = 0.325
    train_crawl_step = 3
    train_subsequence_length = None
    train_predict_two_way = None
    train_freeze_batch_normalization = False
    train_data_pipeline_workers = 8
    train_epochs = 100000
    train_print_frequency = 5000
    train_validate = True
    train_seed = int(round(time.time()))

    # test settings
    test_image_width = 320
    test_image_height = 256
    test_distortion_crop = 0
    test_perform_crop = False
    test_visualize = False            # the reference defaults to True (cv2.imshow); headless boxes cannot
    test_n_measurement_frames = 2
    test_keyframe_buffer_size = 30
    test_keyframe_pose_distance = 0.1
    test_optimal_t_measure = 0.15
    test_optimal_R_measure = 0.0

    dataset = os.environ.get("DVMVS_DATASET", "")
    train_run_directory = os.environ.get("DVMVS_TRAIN_RUNS", os.path.join(_REPO, "training-runs"))
    test_online_scene_path = os.environ.get("DVMVS_SCENE", os.path.join(_REPO, "tests", "golden", "scene000"))
    test_offline_data_path = os.environ.get("DVMVS_DATA", os.path.join(_REPO, "tests", "golden"))
    test_dataset_name = "hololens-dataset"
    test_result_folder = os.environ.get("DVMVS_RESULTS", os.path.join(_REPO, "gpurun_out"))
