"""Write the mp4 of a predicted summary: the kept frames of one video, in order (reference summary.py:1-46).

Passthrough tooling, not part of the accelerated path: it reads ``machine_summary`` of one video from the
``<split>_preds.h5`` file ``predict_dataset`` wrote (h5py when installed, else utils/hdf5.py) and copies the frames whose
entry is 1 from a directory of ``000001.jpg, 000002.jpg, ...`` into ``summary_<video>.mp4`` next to the predictions.
OpenCV is imported only when frames are actually written.
"""
import argparse
import os.path as osp


def read_machine_summary(preds_path, dataset, video):
    """0/1 vector ``preds[dataset][video]["machine_summary"]`` (summary.py:40-43)."""
    from .models import h5_file
    with h5_file(preds_path, "r") as preds:
        return preds[dataset][video]["machine_summary"][...]


def kept_frame_names(summary):
    """File names of the kept frames; frame ``i`` (0-based) lives in ``%06d.jpg % (i + 1)`` (summary.py:14-16)."""
    return [f"{i + 1:06d}.jpg" for i, keep in enumerate(summary) if keep == 1]


def frm2video(frm_dir, summary, vid_writer, width=640, height=480):
    """Append every kept frame, resized to ``width x height``, to ``vid_writer`` (summary.py:11-20)."""
    import cv2
    from tqdm import tqdm
    for name in tqdm(kept_frame_names(summary), ncols=80):
        frame = cv2.imread(osp.join(frm_dir, name))
        vid_writer.write(cv2.resize(frame, (width, height)))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-p", "--path", type=str, required=True, help="Path to hdfs5 predictions file")
    parser.add_argument("-f", "--frames", type=str, required=True, help="Path to frame directory")
    parser.add_argument("-d", "--dataset", type=str, help="Dataset hdfs5 filename")
    parser.add_argument("-v", "--video", type=str, help="Which video key to choose")
    parser.add_argument("--fps", type=int, default=30, help="frames per second")
    parser.add_argument("--width", type=int, default=640, help="frame width")
    parser.add_argument("--height", type=int, default=480, help="frame height")
    args = parser.parse_args(argv)

    summary = read_machine_summary(args.path, args.dataset, args.video)
    import cv2
    summary_path = osp.join(osp.dirname(args.path), f"summary_{args.video}.mp4")
    writer = cv2.VideoWriter(summary_path, cv2.VideoWriter_fourcc(*"mp4v"), args.fps, (args.width, args.height))
    frm2video(args.frames, summary, writer, args.width, args.height)
    writer.release()
    print(f"Summary saved at {summary_path}")


if __name__ == "__main__":
    main()
