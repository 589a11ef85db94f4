"""CLI entry point mirroring the reference's ``main.py`` (main.py:1-13): ``predict`` and
``evaluate``.  ``train`` is outside the accelerated path."""
import typer

from vad_b200.cli import evaluate_vad_from_scratch, predict_vad_from_scratch

app = typer.Typer()


@app.command(name="train")
def train_vad_from_scratch(config_path: str):
    raise typer.BadParameter("training is not part of the B200 inference path; use the reference trainer")


app.command(name="predict")(predict_vad_from_scratch)
app.command(name="evaluate")(evaluate_vad_from_scratch)

if __name__ == "__main__":
    app()
