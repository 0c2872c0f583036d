//! UNBUILT IN THIS ENVIRONMENT.  examples/coin_toss.rs of the reference on the CUDA engine: the same constants,
//! the same builder calls, `CoinTosser` split into its two numeric fields.
use incerto_cuda::*;

const SIMULATION_STEPS: usize = 10_000; // coin_toss.rs:18
const NUM_ENTITIES: usize = 100;        // :19
const ODDS_HEADS: f64 = 0.5;            // :20

fn main() -> Result<(), Error> {
    let mut builder = SimulationBuilder::new();
    let num_heads: Component<u32> = builder.component("CoinTosser.num_heads");
    let num_tosses: Component<u32> = builder.component("CoinTosser.num_tosses");
    let mut simulation = builder
        .add_entity_spawner(move |spawner| {
            // coin_toss.rs:58-63: NUM_ENTITIES x CoinTosser::default()
            spawner.spawn_batch(NUM_ENTITIES as u64, vec![column(num_heads, &[0u32; NUM_ENTITIES]), column(num_tosses, &[0u32; NUM_ENTITIES])]);
        })
        .add_systems(&[System::CoinToss { num_heads, num_tosses, odds_heads: ODDS_HEADS }])
        .build()?;
    simulation.run(SIMULATION_STEPS)?;
    let odds: f64 = simulation.sample_aggregate(&Aggregate::coin_odds(num_heads, num_tosses))?;
    println!("Average odds of heads: {odds}");
    Ok(())
}
